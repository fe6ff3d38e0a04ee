"""CPU checks of the host wrapper module (`aphrodite_engine_b200/_custom_ops.py`): every wrapper keeps the reference's
parameter order / names / defaults (aphrodite/_custom_ops.py) and marshals exactly the arguments its torch op schema
expects. The ops are registered for the CUDA dispatch key only, so calling a wrapper with CPU tensors must travel
through the dispatcher (arity and types accepted) and then fail there — never compute anything on the CPU."""
import inspect

import pytest
import torch

import aphrodite_engine_b200._custom_ops as ops
from aphrodite_engine_b200.scalar_type import scalar_types


def _schema(name):
    fn = getattr(ops, name)
    packet = fn.__globals__["_op"] if "_op" in fn.__globals__ else None
    assert packet is not None
    return packet.default._schema


def _dummy(arg):
    t = str(arg.type)
    if t == "Tensor":
        return torch.zeros(4, 4)
    if t in ("Tensor?", "Optional[Tensor]"):
        return None
    if t in ("Tensor[]", "List[Tensor]"):
        return [torch.zeros(4, 4)]
    if t in ("int", "SymInt"):
        return 1
    if t == "float":
        return 1.0
    if t == "bool":
        return False
    if t == "str":
        return "auto"
    if t in ("int[]", "List[int]"):
        return [1]
    if t in ("str[]", "List[str]"):
        return ["x"]
    if t in ("int[][]", "List[List[int]]"):
        return [[1]]
    if "ScalarType" in t:
        return scalar_types.uint4b8
    raise AssertionError(f"unhandled schema type {t}")


TABLE_NAMES = [n for n in ops.__all__ if n not in ops.COMPOSITE]


@pytest.mark.parametrize("name", TABLE_NAMES)
def test_wrapper_signature_matches_its_op_schema(name):
    sig = inspect.signature(getattr(ops, name))
    schema = _schema(name)
    assert len(sig.parameters) == len(schema.arguments), (name, str(schema))
    # wrappers are positional pass-throughs: no *args / **kwargs, and defaults only at the tail
    kinds = {p.kind for p in sig.parameters.values()}
    assert kinds <= {inspect.Parameter.POSITIONAL_OR_KEYWORD}
    seen_default = False
    for p in sig.parameters.values():
        if p.default is not inspect.Parameter.empty:
            seen_default = True
        else:
            assert not seen_default


def test_reference_defaults_are_kept():
    d = {k: v.default for k, v in inspect.signature(ops.paged_attention_v1).parameters.items()
         if v.default is not inspect.Parameter.empty}
    assert d == {"tp_rank": 0, "blocksparse_local_blocks": 0, "blocksparse_vert_stride": 0,
                 "blocksparse_block_size": 64, "blocksparse_head_sliding_step": 0}
    assert list(inspect.signature(ops.paged_attention_v2).parameters)[:7] == [
        "out", "exp_sum", "max_logits", "tmp_out", "query", "key_cache", "value_cache"]
    d = {k: v.default for k, v in inspect.signature(ops.gptq_marlin_gemm).parameters.items()
         if v.default is not inspect.Parameter.empty}
    assert d == {"has_zp": False, "use_fp32_reduce": False, "is_zp_float": False}
    d = {k: v.default for k, v in inspect.signature(ops.convert_fp8).parameters.items()
         if v.default is not inspect.Parameter.empty}
    assert d == {"scale": 1.0, "kv_dtype": "fp8"}
    assert list(inspect.signature(ops.reshape_and_cache).parameters) == [
        "key", "value", "key_cache", "value_cache", "slot_mapping", "kv_cache_dtype", "k_scale", "v_scale"]
    assert list(inspect.signature(ops.fused_add_rms_norm).parameters) == ["input", "residual", "weight", "epsilon"]


@pytest.mark.parametrize("name", TABLE_NAMES)
def test_wrapper_reaches_the_dispatcher_and_has_no_cpu_path(name):
    schema = _schema(name)
    if not any("Tensor" in str(a.type) for a in schema.arguments):
        pytest.skip("no tensor argument: the op is not dispatched on a device")
    args = [_dummy(a) for a in schema.arguments]
    with pytest.raises((NotImplementedError, RuntimeError)) as ei:
        getattr(ops, name)(*args)
    msg = str(ei.value)
    assert "CPU" in msg or "cuda" in msg.lower() or "GPU" in msg, msg[:300]


def test_composite_wrappers_keep_reference_signatures_and_have_no_cpu_path():
    """The hand-written wrappers (aphrodite/_custom_ops.py:496-513 cutlass_scaled_mm, :632-685 scaled_fp8_quant, the
    sampling pairs): parameter names / defaults of the reference, and CPU tensors end in the dispatcher's error."""
    assert list(inspect.signature(ops.scaled_fp8_quant).parameters) == [
        "input", "scale", "num_token_padding", "scale_ub", "use_per_token_if_dynamic"]
    assert list(inspect.signature(ops.cutlass_scaled_mm).parameters) == ["a", "b", "scale_a", "scale_b", "out_dtype", "bias"]
    assert list(inspect.signature(ops.top_k_top_p_sampling_from_probs).parameters) == [
        "probs", "uniform_samples", "maybe_top_k_arr", "top_k_val", "maybe_top_p_arr", "top_p_val", "deterministic"]
    x = torch.randn(4, 32)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.scaled_fp8_quant(x)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.scaled_fp8_quant(x, use_per_token_if_dynamic=True)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.scaled_fp8_quant(x, torch.ones(1))
    a = torch.zeros(4, 32).to(torch.float8_e4m3fn)
    b = torch.zeros(16, 32).to(torch.float8_e4m3fn).t()
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.cutlass_scaled_mm(a, b, torch.ones(1), torch.ones(1), torch.bfloat16)
    p = torch.full((2, 8), 0.125)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.top_k_sampling_from_probs(p, torch.rand(4, 2), None, 2)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.sampling_from_probs(p, torch.rand(2))
