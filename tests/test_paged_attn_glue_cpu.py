"""CPU checks of the decode attention glue (aphrodite_engine_b200/attention/paged_attn.py) with the ops mocked:
cache views (reference paged_attn.py:49-72), the single-pass / partitioned rule (:127-128: v1 iff max_seq_len <= 8192
and (one 512-token partition or seqs * heads > 512)) and the scratch tensors handed to paged_attention_v2 (:150-165)."""
import pytest
import torch

import aphrodite_engine_b200.attention.paged_attn as pa


class _Recorder:
    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args, **kwargs):
            self.calls.append((name, args, kwargs))
        return fn


@pytest.fixture()
def rec(monkeypatch):
    r = _Recorder()
    monkeypatch.setattr(pa, "ops", r)
    monkeypatch.setattr(pa.PagedAttention, "_ops", r)
    return r


def test_cache_shape_and_views():
    shape = pa.PagedAttention.get_kv_cache_shape(10, 16, 8, 128)
    assert shape == (2, 10, 16 * 8 * 128)
    for dtype, x in ((torch.bfloat16, 8), (torch.uint8, 16), (torch.float32, 4)):
        kv = torch.zeros(shape, dtype=dtype)
        k, v = pa.PagedAttention.split_kv_cache(kv, 8, 128)
        assert k.shape == (10, 8, 128 // x, 16, x) and v.shape == (10, 8, 128, 16)
        assert k.data_ptr() == kv[0].data_ptr() and v.data_ptr() == kv[1].data_ptr()


@pytest.mark.parametrize("seqs,heads,max_len,expect", [
    (1, 8, 512, "v1"),          # one partition
    (1, 8, 513, "v2"),          # two partitions, 8 pairs
    (16, 32, 513, "v2"),        # exactly 512 pairs: not MORE than 512
    (17, 32, 4096, "v1"),       # 544 pairs
    (256, 32, 8192, "v1"),
    (256, 32, 8193, "v2"),      # beyond 8192 cached tokens always partitioned
])
def test_single_pass_versus_partitioned_rule(rec, seqs, heads, max_len, expect):
    D, BS, KV = 64, 16, 4
    kv = torch.zeros(pa.PagedAttention.get_kv_cache_shape(2, BS, KV, D), dtype=torch.float16)
    k, v = pa.PagedAttention.split_kv_cache(kv, KV, D)
    q = torch.zeros(seqs, heads, D, dtype=torch.float16)
    bt = torch.zeros(seqs, 1, dtype=torch.int32)
    sl = torch.ones(seqs, dtype=torch.int32)
    out = pa.PagedAttention.forward_decode(q, k, v, bt, sl, max_len, "auto", KV, 0.125, None, 1.0, 1.0)
    assert out.shape == q.shape and len(rec.calls) == 1
    name, args, _ = rec.calls[0]
    tail = (KV, 0.125, bt, sl, BS, max_len, None, "auto", 1.0, 1.0, 0, 0, 0, 64, 0)
    if expect == "v1":
        assert name == "paged_attention_v1" and args[0] is out and args[1] is q and args[4:] == tail
    else:
        P = -(-max_len // 512)
        assert name == "paged_attention_v2" and args[0] is out and args[4] is q and args[7:] == tail
        exp_sums, max_logits, tmp_out = args[1], args[2], args[3]
        assert exp_sums.shape == (seqs, heads, P) and exp_sums.dtype == torch.float32
        assert max_logits.shape == exp_sums.shape and tmp_out.shape == (seqs, heads, P, D) and tmp_out.dtype == q.dtype


def test_writer_flattens_slots_and_output_buffer_is_reused(rec):
    D, BS, KV = 64, 16, 2
    kv = torch.zeros(pa.PagedAttention.get_kv_cache_shape(2, BS, KV, D), dtype=torch.float16)
    k, v = pa.PagedAttention.split_kv_cache(kv, KV, D)
    key = torch.zeros(3, KV, D, dtype=torch.float16)
    pa.PagedAttention.write_to_paged_cache(key, key, k, v, torch.zeros(3, 1, dtype=torch.int64), "fp8", 0.5, 2.0)
    name, args, _ = rec.calls[0]
    assert name == "reshape_and_cache" and args[4].shape == (3,) and args[5:] == ("fp8", 0.5, 2.0)
    q = torch.zeros(1, 4, D, dtype=torch.float16)
    buf = torch.zeros_like(q)
    got = pa.PagedAttention.forward_decode(q, k, v, torch.zeros(1, 1, dtype=torch.int32), torch.ones(1, dtype=torch.int32),
                                           10, "auto", KV, 1.0, None, 1.0, 1.0, output=buf)
    assert got is buf
    with pytest.raises(AssertionError, match="needs to be a multiple of"):
        pa.PagedAttention.forward_decode(q, k, v, torch.zeros(1, 1, dtype=torch.int32), torch.ones(1, dtype=torch.int32),
                                         10, "auto", KV, 1.0, None, 1.0, 1.0, blocksparse_vert_stride=2,
                                         blocksparse_block_size=24)
