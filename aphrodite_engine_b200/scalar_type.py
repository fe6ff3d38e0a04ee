"""Named sub-byte / float type descriptors (`b_q_type` of gptq_marlin_gemm), as instances of the `_core_C.ScalarType`
custom class — the same names a caller finds in the reference's `aphrodite/scalar_type.py`. Built from one table:
name -> (constructor, arguments); "b<N>" names are GPTQ-style biased types (value = stored - bias)."""
from . import _native

ScalarType = _native.load_core_ext()

_DESCRIPTORS = {
    "int4": ("int_", (4, None)),
    "uint4": ("uint", (4, None)),
    "int8": ("int_", (8, None)),
    "uint8": ("uint", (8, None)),
    "float8_e5m2": ("float_IEEE754", (5, 2)),
    "float16_e8m7": ("float_IEEE754", (8, 7)),
    "float16_e5m10": ("float_IEEE754", (5, 10)),
    "uint4b8": ("uint", (4, 8)),
    "uint8b128": ("uint", (8, 128)),
}
_ALIASES = {"bfloat16": "float16_e8m7", "float16": "float16_e5m10"}


class scalar_types:
    pass


for _n, (_ctor, _args) in _DESCRIPTORS.items():
    setattr(scalar_types, _n, getattr(ScalarType, _ctor)(*_args))
for _n, _target in _ALIASES.items():
    setattr(scalar_types, _n, getattr(scalar_types, _target))
del _n, _ctor, _args, _target
