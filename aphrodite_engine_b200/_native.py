"""Loads the two native artefacts and fails loudly when they are missing.

  libb200decode.so  — C ABI (include/b200_decode.h); exposed here through ctypes as `lib`
  _C.abi3.so        — TORCH_LIBRARY(_C, _C_cache_ops, _C_cuda_utils) shim -> torch.ops._C.*

Nothing in this module (or anything it imports) touches oracle/: the product path is CUDA-only.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libb200decode.so")
SHIM_PATH = os.path.join(_PKG, "_C.abi3.so")
CORE_PATH = os.path.join(_PKG, "_core_C.abi3.so")
MOE_PATH = os.path.join(_PKG, "_moe_C.abi3.so")

_lib = None
_ops_loaded = False

c_void_p, c_int, c_int64, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argtypes, mirroring include/b200_decode.h (restype int unless listed in _RESTYPES)
SIGNATURES = {
    "b200_last_error": [],
    "b200_abi_version": [],
    "b200_parse_kv_cache_dtype": [ctypes.c_char_p],
    "b200_paged_attention_v1": [c_void_p] * 4 + [c_int] * 5 + [c_float] + [c_void_p] * 2 +
                               [c_int] * 2 + [c_void_p] + [c_int64] * 3 + [c_int] * 2 +
                               [c_float] * 2 + [c_int] * 5 + [c_void_p],
    "b200_paged_attention_v2": [c_void_p] * 7 + [c_int] * 5 + [c_float] + [c_void_p] * 2 +
                               [c_int] * 3 + [c_void_p] + [c_int64] * 3 + [c_int] * 2 +
                               [c_float] * 2 + [c_int] * 5 + [c_void_p],
    "b200_set_attention_impl": [c_int],
    "b200_last_attention_path": [],
    "b200_last_attention_cluster_split": [],
    "b200_reshape_and_cache": [c_void_p] * 5 + [c_int] * 5 + [c_int64] * 2 + [c_int] * 2 +
                              [c_float] * 2 + [c_void_p],
    "b200_reshape_and_cache_flash": [c_void_p] * 5 + [c_int] * 4 + [c_int64] * 3 + [c_int] * 2 +
                                    [c_float] * 2 + [c_void_p],
    "b200_copy_blocks": [c_void_p] * 3 + [c_int] * 2 + [c_int64, c_void_p],
    "b200_swap_blocks": [c_void_p] * 3 + [c_int, c_int64, c_int, c_void_p],
    "b200_convert_fp8": [c_void_p] * 2 + [c_int64] + [c_int] * 3 + [c_float, c_void_p],
    "b200_rms_norm": [c_void_p] * 3 + [c_float] + [c_int] * 3 + [c_void_p],
    "b200_fused_add_rms_norm": [c_void_p] * 3 + [c_float] + [c_int] * 3 + [c_void_p],
    "b200_rotary_embedding": [c_void_p] * 5 + [c_int] * 5 + [c_int64] * 2 + [c_int] * 2 + [c_void_p],
    "b200_rotary_embedding_and_cache": [c_void_p] * 8 + [c_int] * 5 + [c_int64] * 3 + [c_int] * 5 + [c_float] * 2 +
                                       [c_void_p],
    "b200_act_and_mul": [c_void_p] * 2 + [c_int] * 4 + [c_void_p],
    "b200_activation": [c_void_p] * 2 + [c_int] * 4 + [c_void_p],
    "b200_marlin_gemm_plan": [c_int] * 4,
    "b200_debug_marlin_prof": [c_void_p],
    "b200_gptq_marlin_gemm": [c_void_p] * 7 + [c_int] * 8 + [c_void_p],
    "b200_marlin_gemm_moe": [c_void_p] * 3 + [c_int64] + [c_void_p] * 7 + [c_int] * 10 + [c_void_p],
    "b200_gptq_marlin_repack": [c_void_p] * 3 + [c_int] * 3 + [c_void_p],
    "b200_awq_marlin_repack": [c_void_p] * 2 + [c_int] * 3 + [c_void_p],
    "b200_moe_align_block_size": [c_void_p, c_int, c_int64, c_int, c_int] + [c_void_p] * 4,
    "b200_topk_softmax": [c_void_p] * 4 + [c_int] * 3 + [c_void_p],
    "b200_moe_expert_scale_add": [c_void_p] * 4 + [c_int] * 6 + [c_void_p],
    "b200_permute_cols": [c_void_p] * 3 + [c_int64, c_int, c_void_p],
    "b200_awq_dequantize": [c_void_p] * 4 + [c_int64, c_int, c_int, c_void_p],
    "b200_advance_step_flashattn": [c_int] * 3 + [c_void_p] * 6 + [c_int64, c_void_p],
    "b200_car_meta_size": [],
    "b200_car_init": [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int],
    "b200_car_dispose": [c_int64],
    "b200_car_register_buffer": [c_int64, c_void_p, c_void_p, c_void_p],
    "b200_car_all_reduce": [c_int64, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "b200_car_get_graph_buffer_ipc_meta": [c_int64, c_void_p, c_void_p, c_int],
    "b200_car_register_graph_buffers": [c_int64, c_void_p, c_void_p, c_int],
    "b200_tp_flag_bytes": [],
    "b200_tp_set_stamp_buffer": [c_void_p],
    "b200_tp_allreduce_rows": [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                               c_float] + [c_int] * 6 + [c_void_p],
    "b200_get_device_attribute": [c_int64, c_int64],
    "b200_get_max_shared_memory_per_block_device_attribute": [c_int64],
    # fp8 activation quantisation, W8A8 GEMM, sampling, prefill attention over the paged cache
    "b200_static_scaled_fp8_quant": [c_void_p] * 3 + [c_int64, c_int, c_void_p],
    "b200_dynamic_scaled_fp8_quant": [c_void_p] * 3 + [c_int64, c_int, c_void_p],
    "b200_dynamic_per_token_scaled_fp8_quant": [c_void_p] * 4 + [c_int] * 3 + [c_void_p],
    "b200_cutlass_scaled_mm_supports_fp8": [c_int],
    "b200_scaled_mm_plan": [c_int] * 3,
    "b200_scaled_mm_set_tile": [c_int],
    "b200_cutlass_scaled_mm": [c_void_p] * 6 + [c_int] * 3 + [c_int64] * 3 + [c_int] * 5 + [c_void_p] * 2,
    "b200_sampling_from_probs": [c_void_p] * 3 + [c_int] * 3 + [c_void_p],
    "b200_rejection_sampling_from_probs": [c_int] + [c_void_p] * 5 + [c_int, c_void_p, c_float] + [c_int] * 4 +
                                          [c_void_p],
    "b200_top_p_renorm_prob": [c_void_p] * 3 + [c_float] + [c_int] * 2 + [c_void_p],
    "b200_top_k_renorm_prob": [c_void_p] * 3 + [c_int] * 3 + [c_void_p],
    "b200_top_k_mask_logits": [c_void_p] * 3 + [c_int] * 3 + [c_void_p],
    "b200_context_attention_fwd": [c_void_p] * 11 + [c_int] * 7 + [c_int64] * 13 + [c_float] * 3 + [c_int] * 3 +
                                  [c_void_p],
}
_RESTYPES = {
    "b200_tp_flag_bytes": c_int64,
    "b200_tp_set_stamp_buffer": None,
    "b200_car_meta_size": c_int64,
    "b200_car_init": c_int64,
    "b200_car_dispose": None,
    "b200_last_error": ctypes.c_char_p,
    "b200_get_device_attribute": c_int64,
    "b200_get_max_shared_memory_per_block_device_attribute": c_int64,
}


class NativeLibraryMissing(ImportError):
    pass


def _missing(path):
    return NativeLibraryMissing(
        f"{path} is missing: the B200 decode path has no CPU/PyTorch fallback. Build it with "
        f"`python -c 'import __graft_entry__ as g; g.build()'` (nvcc -gencode arch=compute_100a,code=sm_100a).")


def load_c_abi():
    """ctypes handle on libb200decode.so with argtypes set for every symbol of the header."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise _missing(LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def load_torch_ops():
    """Registers torch.ops._C.* / _C_cache_ops.* / _C_cuda_utils.* from the in-tree shim."""
    global _ops_loaded
    if not _ops_loaded:
        if not os.path.exists(SHIM_PATH):
            raise _missing(SHIM_PATH)
        load_c_abi()
        load_core_ext()
        torch.ops.load_library(SHIM_PATH)
        if not os.path.exists(MOE_PATH):
            raise _missing(MOE_PATH)
        torch.ops.load_library(MOE_PATH)
        _ops_loaded = True
    return torch.ops._C


def load_core_ext():
    """Makes `torch.classes._core_C.ScalarType` available (the `b_q_type` argument type of the marlin op).
    If the reference's own `_core_C` extension already registered the class, it is used as is."""
    try:
        return torch.classes._core_C.ScalarType
    except Exception:
        pass
    if not os.path.exists(CORE_PATH):
        raise _missing(CORE_PATH)
    torch.ops.load_library(CORE_PATH)
    return torch.classes._core_C.ScalarType


def last_error() -> str:
    return load_c_abi().b200_last_error().decode()
