#!/usr/bin/env python3
"""Generates tests/golden/marlin_*.npz by IMPORTING THE REFERENCE'S OWN PYTHON (build container only):
aphrodite.quantization.utils.{quant_utils,marlin_utils,marlin_utils_test} from /root/reference, loaded
through a stub `aphrodite` package because the full package does not import here (triton / librosa
version drift). The vectors pin oracle/marlin.py and, on the GPU, the repack and GEMM kernels.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import build_ref
    core = build_ref.build_core_ext()          # the reference's own _core_C (ScalarType), built with g++
    pkg = types.ModuleType("aphrodite")
    pkg.__path__ = [os.path.join(REF, "aphrodite"), os.path.dirname(core)]
    sys.modules["aphrodite"] = pkg
    tu = types.ModuleType("aphrodite.triton_utils")
    tu.HAS_TRITON = False
    sys.modules["aphrodite.triton_utils"] = tu
    try:
        import loguru  # noqa: F401
    except ImportError:
        lg = types.ModuleType("loguru")

        class _L:
            def __getattr__(self, k):
                return lambda *a, **kw: None
        lg.logger = _L()
        sys.modules["loguru"] = lg
    from aphrodite.quantization.utils import marlin_utils, marlin_utils_test, quant_utils
    from aphrodite.scalar_type import scalar_types
    return marlin_utils, marlin_utils_test, quant_utils, scalar_types


def main():
    mu, mut, qu, st = import_reference()
    for tag, dt in (("bf16", torch.bfloat16), ("f16", torch.float16)):
        for bits, qt_sym, qt_zp in ((4, st.uint4b8, st.uint4), (8, st.uint8b128, st.uint8)):
            for gs in (128, -1, 32):
                if tag == "f16" and not (bits == 4 and gs == 128):
                    continue            # one fp16 case is enough: only the rounding dtype differs
                torch.manual_seed(bits * 1000 + (gs % 997))
                K, N = 256, 128
                w = torch.randn(K, N).to(dt)
                w_ref, mq, ms, g_idx, sort_idx, _ = mut.marlin_quantize(w, qt_sym, gs, False)
                _, q_w, s, _ = qu.quantize_weights(w, qt_sym, gs)
                gp = qu.gptq_pack(q_w, bits, K, N)
                d = dict(w=w.float().numpy(), w_ref=w_ref.float().numpy(), marlin_q_w=mq.numpy(),
                         marlin_s=ms.float().numpy(), q_w=q_w.numpy(), s=s.float().numpy(),
                         gptq_packed=gp.numpy())
                if gs != -1 and gs != 32:
                    w_ref2, mq2, ms2, mzp = mut.awq_marlin_quantize(w, qt_zp, gs)
                    _, q_w2, s2, zp2 = qu.quantize_weights(w, qt_zp, gs, zero_points=True)
                    ap = qu.awq_pack(q_w2, bits, K, N)
                    azp = qu.awq_pack(zp2, bits, K // gs, N)
                    d.update(awq_w_ref=w_ref2.float().numpy(), awq_marlin_q_w=mq2.numpy(),
                             awq_marlin_s=ms2.float().numpy(), awq_marlin_zp=mzp.numpy(),
                             awq_q_w=q_w2.numpy(), awq_zp=zp2.numpy(), awq_packed=ap.numpy(),
                             awq_zp_packed=azp.numpy(),
                             awq_zp_to_marlin=mu.awq_to_marlin_zero_points(azp, K // gs, N, bits).numpy())
                name = f"marlin_{tag}_b{bits}_g{gs if gs != -1 else 'ch'}"
                np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
                print("wrote", name)


if __name__ == "__main__":
    main()
