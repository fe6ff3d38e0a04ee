"""CPU-only: the C-ABI library loads without a GPU and exports every symbol include/*.h declares;
the torch-op shim registers the reference's schemas (kernels/torch_bindings.cpp) verbatim; the
product package never imports the oracle. No compute calls here."""
import ctypes
import glob
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(b200_\w+)\s*\(", txt))
    return syms


def test_header_symbols_are_exported_and_bound():
    from aphrodite_engine_b200 import _native
    lib = _native.load_c_abi()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported by libb200decode.so"
    assert declared == set(_native.SIGNATURES), "ctypes table and header disagree"
    assert lib.b200_abi_version() == 1
    assert lib.b200_parse_kv_cache_dtype(b"auto") == 0
    assert lib.b200_parse_kv_cache_dtype(b"fp8") == 1 == lib.b200_parse_kv_cache_dtype(b"fp8_e4m3")
    assert lib.b200_parse_kv_cache_dtype(b"fp8_e5m2") == 2
    assert lib.b200_parse_kv_cache_dtype(b"int8") == -1


def test_library_is_sm100a_only_and_has_no_torch_dependency():
    from aphrodite_engine_b200 import _native
    out = subprocess.run(["ldd", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if os.path.exists(cuobjdump):
        arch = subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
        assert "sm_100a" in arch
        assert not re.search(r"sm_(?!100a)\d+", arch), arch


# Schema strings copied from the reference's registration file (kernels/torch_bindings.cpp:25-124,
# :456-504): the drop-in must register byte-for-byte equivalent signatures (names, types, `!` marks).
REF_SCHEMAS = {
    "_C::paged_attention_v1":
        "paged_attention_v1(Tensor! out, Tensor query, Tensor key_cache, Tensor value_cache, int num_kv_heads,"
        " float scale, Tensor block_tables, Tensor seq_lens, int block_size, int max_seq_len,"
        " Tensor? alibi_slopes, str kv_cache_dtype, float k_scale, float v_scale, int tp_rank,"
        " int blocksparse_local_blocks, int blocksparse_vert_stride, int blocksparse_block_size,"
        " int blocksparse_head_sliding_step) -> ()",
    "_C::paged_attention_v2":
        "paged_attention_v2(Tensor! out, Tensor! exp_sums, Tensor! max_logits, Tensor! tmp_out, Tensor query,"
        " Tensor key_cache, Tensor value_cache, int num_kv_heads, float scale, Tensor block_tables,"
        " Tensor seq_lens, int block_size, int max_seq_len, Tensor? alibi_slopes, str kv_cache_dtype,"
        " float k_scale, float v_scale, int tp_rank, int blocksparse_local_blocks,"
        " int blocksparse_vert_stride, int blocksparse_block_size, int blocksparse_head_sliding_step) -> ()",
    "_C::silu_and_mul": "silu_and_mul(Tensor! out, Tensor input) -> ()",
    "_C::gelu_and_mul": "gelu_and_mul(Tensor! out, Tensor input) -> ()",
    "_C::gelu_tanh_and_mul": "gelu_tanh_and_mul(Tensor! out, Tensor input) -> ()",
    "_C::gelu_new": "gelu_new(Tensor! out, Tensor input) -> ()",
    "_C::gelu_fast": "gelu_fast(Tensor! out, Tensor input) -> ()",
    "_C::gelu_quick": "gelu_quick(Tensor! out, Tensor input) -> ()",
    "_C::rms_norm": "rms_norm(Tensor! out, Tensor input, Tensor weight, float epsilon) -> ()",
    "_C::fused_add_rms_norm":
        "fused_add_rms_norm(Tensor! input, Tensor! residual, Tensor weight, float epsilon) -> ()",
    "_C::rotary_embedding":
        "rotary_embedding(Tensor positions, Tensor! query, Tensor! key, int head_size, Tensor cos_sin_cache,"
        " bool is_neox) -> ()",
    "_C::batched_rotary_embedding":
        "batched_rotary_embedding(Tensor positions, Tensor! query, Tensor! key, int head_size,"
        " Tensor cos_sin_cache, bool is_neox, int rot_dim, Tensor cos_sin_cache_offsets) -> ()",
    "_C_cache_ops::swap_blocks": "swap_blocks(Tensor src, Tensor! dst, Tensor block_mapping) -> ()",
    "_C_cache_ops::copy_blocks":
        "copy_blocks(Tensor(a!)[] key_caches, Tensor[](b!) value_caches, Tensor block_mapping) -> ()",
    "_C_cache_ops::reshape_and_cache":
        "reshape_and_cache(Tensor key, Tensor value, Tensor! key_cache, Tensor! value_cache,"
        " Tensor slot_mapping, str kv_cache_dtype, float k_scale, float v_scale) -> ()",
    "_C_cache_ops::reshape_and_cache_flash":
        "reshape_and_cache_flash(Tensor key, Tensor value, Tensor! key_cache, Tensor! value_cache,"
        " Tensor slot_mapping, str kv_cache_dtype, float k_scale, float v_scale) -> ()",
    "_C_cache_ops::convert_fp8":
        "convert_fp8(Tensor! dst_cache, Tensor src_cache, float scale, str kv_cache_dtype) -> ()",
    "_C::gptq_marlin_gemm":
        "gptq_marlin_gemm(Tensor a, Tensor b_q_weight, Tensor b_scales, Tensor b_zeros, Tensor g_idx, Tensor perm,"
        " Tensor workspace, __torch__.torch.classes._core_C.ScalarType b_q_type, int size_m, int size_n,"
        " int size_k, bool is_k_full, bool has_zp, bool use_fp32_reduce, bool is_zp_float) -> Tensor",
    "_C::gptq_marlin_repack":
        "gptq_marlin_repack(Tensor b_q_weight, Tensor perm, SymInt size_k, SymInt size_n, int num_bits) -> Tensor",
    "_C::awq_marlin_repack":
        "awq_marlin_repack(Tensor b_q_weight, SymInt size_k, SymInt size_n, int num_bits) -> Tensor",
    "_C::moe_align_block_size":
        "moe_align_block_size(Tensor topk_ids, int num_experts, int block_size, Tensor! sorted_token_ids,"
        " Tensor! experts_ids, Tensor! num_tokens_post_pad) -> ()",
    "_moe_C::topk_softmax":
        "topk_softmax(Tensor! topk_weights, Tensor! topk_indices, Tensor! token_expert_indices,"
        " Tensor gating_output) -> ()",
    "_C::advance_step_flashattn":
        "advance_step_flashattn(int num_seqs, int num_queries, int block_size, Tensor! input_tokens,"
        " Tensor sampled_token_ids, Tensor! input_positions, Tensor! seq_lens, Tensor! slot_mapping,"
        " Tensor block_tables) -> ()",
    "_C::awq_dequantize":
        "awq_dequantize(Tensor _kernel, Tensor _scaling_factors, Tensor _zeros, int split_k_iters, int thx,"
        " int thy) -> Tensor",
    "_C::permute_cols": "permute_cols(Tensor A, Tensor perm) -> Tensor",
    "_C_custom_ar::init_custom_ar":
        "init_custom_ar(Tensor meta, Tensor rank_data, str[] handles, int[] offsets, int rank,"
        " bool full_nvlink) -> int",
    "_C_custom_ar::all_reduce_reg": "all_reduce_reg(int fa, Tensor inp, Tensor! out) -> ()",
    "_C_custom_ar::all_reduce_unreg": "all_reduce_unreg(int fa, Tensor inp, Tensor reg_buffer, Tensor! out) -> ()",
    "_C_custom_ar::register_buffer": "register_buffer(int fa, Tensor t, str[] handles, int[] offsets) -> ()",
    # round 2, SURVEY §8(f): kernels/torch_bindings.cpp:235-253 (W8A8 GEMM), :294-350 (sampling), :374-390 (fp8 quant)
    "_C::cutlass_scaled_mm":
        "cutlass_scaled_mm(Tensor! out, Tensor a, Tensor b, Tensor a_scales, Tensor b_scales, Tensor? bias) -> ()",
    "_C::cutlass_scaled_mm_supports_fp8": "cutlass_scaled_mm_supports_fp8(int cuda_device_capability) -> bool",
    "_C::cutlass_scaled_mm_azp":
        "cutlass_scaled_mm_azp(Tensor! out, Tensor a, Tensor b, Tensor a_scales, Tensor b_scales, Tensor azp_adj,"
        " Tensor? azp, Tensor? bias) -> ()",
    "_C::sampling_from_probs": "sampling_from_probs(Tensor probs, Tensor uniform_samples, bool deterministic) -> Tensor",
    "_C::top_k_sampling_from_probs":
        "top_k_sampling_from_probs(Tensor probs, Tensor uniform_samples, Tensor? maybe_top_k_arr, int top_k_val,"
        " bool deterministic) -> Tensor[]",
    "_C::min_p_sampling_from_probs":
        "min_p_sampling_from_probs(Tensor probs, Tensor uniform_samples, Tensor? maybe_min_p_arr, float min_p_val,"
        " bool deterministic) -> Tensor[]",
    "_C::top_p_sampling_from_probs":
        "top_p_sampling_from_probs(Tensor probs, Tensor uniform_samples, Tensor? maybe_top_p_arr, float top_p_val,"
        " bool deterministic) -> Tensor[]",
    "_C::top_k_top_p_sampling_from_probs":
        "top_k_top_p_sampling_from_probs(Tensor probs, Tensor uniform_samples, Tensor? maybe_top_k_arr,"
        " float top_k_val, Tensor? maybe_top_p_arr, float top_p_val, bool deterministic) -> Tensor[]",
    "_C::top_k_renorm_prob": "top_k_renorm_prob(Tensor probs, Tensor? maybe_top_k_arr, int top_k_val) -> Tensor",
    "_C::top_p_renorm_prob": "top_p_renorm_prob(Tensor probs, Tensor? maybe_top_p_arr, float top_p_val) -> Tensor",
    "_C::top_k_mask_logits": "top_k_mask_logits(Tensor logits, Tensor? maybe_top_k_arr, int top_k_val) -> Tensor",
    "_C::static_scaled_fp8_quant": "static_scaled_fp8_quant(Tensor! out, Tensor input, Tensor scale) -> ()",
    "_C::dynamic_scaled_fp8_quant": "dynamic_scaled_fp8_quant(Tensor! out, Tensor input, Tensor! scale) -> ()",
    "_C::dynamic_per_token_scaled_fp8_quant":
        "dynamic_per_token_scaled_fp8_quant(Tensor! out, Tensor input, Tensor! scale, Tensor? scale_ub) -> ()",
    "_C_cuda_utils::get_device_attribute": "get_device_attribute(int attribute, int device_id) -> int",
    "_C_cuda_utils::get_max_shared_memory_per_block_device_attribute":
        "get_max_shared_memory_per_block_device_attribute(int device_id) -> int",
}


def _sig(schema):
    return ([(a.name, str(a.type), bool(a.alias_info and a.alias_info.is_write)) for a in schema.arguments],
            [str(r.type) for r in schema.returns])


@pytest.mark.parametrize("op", sorted(REF_SCHEMAS))
def test_torch_ops_registered_with_reference_schema(op):
    from aphrodite_engine_b200 import _native
    _native.load_torch_ops()
    ns, name = op.split("::")
    packet = getattr(getattr(torch.ops, ns), name)
    ours = packet.default._schema
    ref = torch._C.parse_schema(f"{ns}::{REF_SCHEMAS[op]}")
    assert _sig(ours) == _sig(ref), f"{op}: {ours} != {ref}"


def test_custom_ar_inferred_schema_ops_exist():
    from aphrodite_engine_b200 import _native
    _native.load_torch_ops()
    for name in ("dispose", "meta_size", "get_graph_buffer_ipc_meta", "register_graph_buffers"):
        assert hasattr(torch.ops._C_custom_ar, name)
    assert torch.ops._C_custom_ar.meta_size() > 0 and torch.ops._C_custom_ar.meta_size() % 128 == 0


def test_shim_exports_pyinit_for_import_as_aphrodite_C():
    from aphrodite_engine_b200 import _native
    lib = ctypes.CDLL(_native.SHIM_PATH)
    assert hasattr(lib, "PyInit__C")
    assert hasattr(ctypes.CDLL(_native.MOE_PATH), "PyInit__moe_C")
    assert hasattr(ctypes.CDLL(_native.CORE_PATH), "PyInit__core_C")


def test_scalar_type_class_mirrors_reference_interface():
    from aphrodite_engine_b200.scalar_type import ScalarType, scalar_types as st
    assert (st.uint4b8.size_bits, st.uint4b8.bias, st.uint4b8.min(), st.uint4b8.max()) == (4, 8, -8, 7)
    assert (st.uint4.min(), st.uint4.max(), st.uint8b128.min(), st.uint8b128.max()) == (0, 15, -128, 127)
    assert st.int8.min() == -128 and st.int8.max() == 127 and st.int8.is_signed()
    assert st.uint4b8.is_integer() and not st.uint4b8.is_floating_point() and st.uint4b8.has_bias()
    assert str(st.uint4b8) == "uint4b8" and repr(st.uint4) == "ScalarType.uint4"
    assert st.float16_e5m10.max() == 65504.0 and st.float16_e5m10.is_ieee_754()
    assert st.uint4b8 == ScalarType.uint(4, 8) and not (st.uint4b8 == st.uint4)
    assert ScalarType(0, 4, 8, False) == st.uint4b8


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "aphrodite_engine_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/" in txt and f.endswith(".py") \
                        and re.search(r"(open|load|CDLL)\(.*oracle", txt):
                    offenders.append(f)
    assert not offenders, offenders


def test_missing_extension_fails_loudly(tmp_path, monkeypatch):
    from aphrodite_engine_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU/PyTorch fallback"):
        _native.load_c_abi()


def test_ops_refuse_cpu_tensors():
    """There is no CPU dispatch key registered: a CPU call must raise, not silently compute."""
    import aphrodite_engine_b200._custom_ops as ops
    x = torch.randn(2, 8)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.silu_and_mul(torch.empty(2, 4), x)


def test_header_is_plain_c_and_every_symbol_links(tmp_path):
    """include/b200_decode.h is the drop-in boundary for NON-torch hosts (cgo / JNI / ctypes stubs): it must compile as
    plain C99 and every declared function must resolve against libb200decode.so at link time."""
    from aphrodite_engine_b200 import _native
    syms = sorted(_declared_symbols())
    src = tmp_path / "link_all.c"
    body = "\n".join(f"    p[{i}] = (fn_t)&{s};" for i, s in enumerate(syms))
    src.write_text('#include "b200_decode.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n'
                   f"    fn_t p[{len(syms)}];\n{body}\n"
                   f'    printf("%d %d\\n", b200_abi_version(), (int)(sizeof p / sizeof p[0]));\n    return p[0] == 0;\n}}\n')
    exe = tmp_path / "link_all"
    libdir = os.path.dirname(_native.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{os.path.join(ROOT, 'include')}", str(src),
                        "-o", str(exe), f"-L{libdir}", "-lb200decode", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["1", str(len(syms))], (out.stdout, out.stderr)
