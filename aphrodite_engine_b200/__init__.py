"""aphrodite_engine_b200 — the sm_100a decode hot path behind aphrodite-engine's custom-op boundary.

Only what the path needs lives here:
  csrc/            hand-written CUDA kernels + the C ABI (include/b200_decode.h) + the torch-op shim
  _native.py       loader for libb200decode.so (ctypes, C ABI) and _C.abi3.so (torch.ops._C.*)
  _custom_ops.py   host-side mirror of aphrodite/_custom_ops.py for the in-scope ops
  attention/       mirror of aphrodite/attention/ops/paged_attn.py (V1/V2 heuristic, cache views)

There is NO CPU fallback: importing `_custom_ops` without the built CUDA extension raises.
"""
__version__ = "0.1.0"
