"""CPU-only: pins oracle/marlin.py (restatement of the Marlin weight formats) against vectors produced
by the reference's own Python utilities (tests/golden/marlin_*.npz, make_golden_marlin.py). Integer
layouts are compared bit-exactly."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import marlin as om
from tests.golden_io import GOLDEN_DIR

FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "marlin_*.npz")))


def _load(path):
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    _, tag, b, g = name.split("_")
    dt = torch.bfloat16 if tag == "bf16" else torch.float16
    gs = -1 if g == "gch" else int(g[1:])
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}, dt, int(b[1:]), gs


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p)[:-4] for p in FILES])
def test_marlin_formats_match_reference_python(path):
    g, dt, bits, gs = _load(path)
    K, N = g["w"].shape
    w = g["w"].to(dt)
    w_ref, mq, ms = om.marlin_quantize(w, bits, gs)
    assert torch.equal(w_ref.float(), g["w_ref"])
    assert torch.equal(mq, g["marlin_q_w"])                   # tile permutation + packing: bit-exact
    assert torch.equal(ms.float(), g["marlin_s"])
    _, q_w, s, _ = om.quantize_weights(w, bits, gs, 1 << (bits - 1))
    assert torch.equal(q_w, g["q_w"]) and torch.equal(s.float(), g["s"])
    gp = om.pack_rows(q_w, bits)
    assert torch.equal(gp, g["gptq_packed"])
    assert torch.equal(om.gptq_marlin_repack(gp, None, K, N, bits), g["marlin_q_w"])
    # act-order style row gather before tiling
    perm = torch.randperm(K, generator=torch.Generator().manual_seed(1)).int()
    assert torch.equal(om.gptq_marlin_repack(gp, perm, K, N, bits), om.marlin_weights(q_w[perm.long()], bits))
    if "awq_packed" in g:
        w_ref2, mq2, ms2, mzp = om.awq_marlin_quantize(w, bits, gs)
        assert torch.equal(w_ref2.float(), g["awq_w_ref"])
        assert torch.equal(mq2, g["awq_marlin_q_w"]) and torch.equal(ms2.float(), g["awq_marlin_s"])
        assert torch.equal(mzp, g["awq_marlin_zp"])
        assert torch.equal(om.awq_pack(g["awq_q_w"], bits), g["awq_packed"])
        assert torch.equal(om.awq_marlin_repack(g["awq_packed"], K, N, bits), g["awq_marlin_q_w"])
        assert torch.equal(om.awq_to_marlin_zero_points(g["awq_zp_packed"], K // gs, N, bits),
                           g["awq_zp_to_marlin"])


def test_weight_perm_is_a_permutation_of_a_16x64_block():
    for bits in (4, 8):
        p = om.weight_perm(bits)
        assert p.numel() == 1024 and torch.equal(torch.sort(p).values, torch.arange(1024))
