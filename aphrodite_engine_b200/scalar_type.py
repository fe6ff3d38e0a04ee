"""Mirror of aphrodite/scalar_type.py: named instances of the `_core_C.ScalarType` custom class
(sub-byte dtype descriptors used as the `b_q_type` argument of gptq_marlin_gemm)."""
from . import _native

ScalarType = _native.load_core_ext()


class scalar_types:
    int4 = ScalarType.int_(4, None)
    uint4 = ScalarType.uint(4, None)
    int8 = ScalarType.int_(8, None)
    uint8 = ScalarType.uint(8, None)
    float8_e5m2 = ScalarType.float_IEEE754(5, 2)
    float16_e8m7 = ScalarType.float_IEEE754(8, 7)
    float16_e5m10 = ScalarType.float_IEEE754(5, 10)
    # "gptq" types: value = stored - bias
    uint4b8 = ScalarType.uint(4, 8)
    uint8b128 = ScalarType.uint(8, 128)
    bfloat16 = float16_e8m7
    float16 = float16_e5m10
