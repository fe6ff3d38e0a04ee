#!/usr/bin/env python3
"""Build the reference's OWN CUDA kernels for sm_100a into oracle/_ref/_ref_cuda_C.so (test infrastructure only).

TEST INFRASTRUCTURE — never imported by the product path. Used by tests/test_gpu_vs_ref_cuda.py (same-process,
same-tensor parity of this repo's kernels against the reference's kernels on the B200) and by
tests/bench_vs_ref_cuda.py (the "reference kernel recompiled for sm_100a" column of SURVEY §8d).

Compiles, *where they lie* under /root/reference/kernels (no sources are copied into this repo), the hot-path
translation units with nvcc directly — not the reference's CMake build, which pins arch lists without sm_100
(CMakeLists.txt:23) and fetches CUTLASS from the network (:231-241):

    attention/attention_kernels.cu  cache_kernels.cu  layernorm_kernels.cu  pos_encoding_kernels.cu
    activation_kernels.cu  permute_cols.cu  prepare_inputs/advance_step.cu
    quantization/gptq_marlin/{gptq_marlin,gptq_marlin_repack,awq_marlin_repack}.cu  quantization/awq/gemm_kernels.cu
    moe/{align_block_size_kernel,softmax,marlin_moe_ops}.cu  all_reduce/custom_all_reduce.cu
    quantization/fp8/common.cu  sampling/sampling.cu

with the flags of cmake/utils.cmake:93-111 (torch's COMMON_NVCC_FLAGS minus the __CUDA_NO_HALF* set, -DENABLE_FP8)
and `-gencode arch=compute_100a,code=sm_100a`. Registration: oracle/ref_cuda_bindings.cpp (namespace _ref_cuda_C).

The GPU box has no /root/reference: it uses the prebuilt .so (git-ignored, shipped by gpurun). No-op there.
"""
import os
import subprocess
import sys
import sysconfig
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("APHRODITE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
TARGET = os.path.join(OUT, "_ref_cuda_C.so")

CU_SRCS = [
    "attention/attention_kernels.cu",
    "cache_kernels.cu",
    "layernorm_kernels.cu",
    "pos_encoding_kernels.cu",
    "activation_kernels.cu",
    "permute_cols.cu",
    "prepare_inputs/advance_step.cu",
    "quantization/gptq_marlin/gptq_marlin.cu",
    "quantization/gptq_marlin/gptq_marlin_repack.cu",
    "quantization/gptq_marlin/awq_marlin_repack.cu",
    "quantization/awq/gemm_kernels.cu",
    "moe/align_block_size_kernel.cu",
    "moe/softmax.cu",
    "moe/marlin_moe_ops.cu",
    "quantization/fp8/common.cu",        # static / dynamic / per-token scaled_fp8_quant (round 2: §8 f4)
    "sampling/sampling.cu",              # sampling_from_probs, top-k / top-p / min-p samplers, renorm, mask (§8 f2)
]
# the reference's own TP all-reduce (the N > 1 legs of bench.py's ref_cuda arm): its own library, it needs libcuda
AR_SRC = "all_reduce/custom_all_reduce.cu"
AR_TARGET = os.path.join(OUT, "_ref_cuda_ar_C.so")


def _run(cmd):
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-4000:] + r.stderr[-8000:])
        raise RuntimeError("reference CUDA build failed: " + cmd[-3])
    return time.time() - t0


def build(force: bool = False) -> bool:
    """Returns True if oracle/_ref/_ref_cuda_C.so exists afterwards."""
    bind_src = os.path.join(HERE, "ref_cuda_bindings.cpp")
    ar_bind_src = os.path.join(HERE, "ref_cuda_ar_bindings.cpp")
    fresh = (os.path.exists(TARGET) and os.path.getmtime(TARGET) >= os.path.getmtime(bind_src) and
             os.path.exists(AR_TARGET) and os.path.getmtime(AR_TARGET) >= os.path.getmtime(ar_bind_src))
    if fresh and not force:
        return True
    if not os.path.isdir(os.path.join(REF, "kernels", "attention")):
        return os.path.exists(TARGET)
    import torch

    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    pyinc = sysconfig.get_paths()["include"]
    nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    inc = [f"-I{REF}/kernels", f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include", f"-I{pyinc}",
           "-I/usr/local/cuda/include"]
    defs = ["-DTORCH_EXTENSION_NAME=_ref_cuda_C", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DENABLE_FP8"]
    cu = [nvcc, "-std=c++17", "-O2", "-w", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
          "--threads", "4", "-Xcompiler", "-fPIC", *defs, *inc]
    objdir = os.path.join(OUT, "obj_cuda")
    os.makedirs(objdir, exist_ok=True)
    jobs, objs = [], []
    for s in CU_SRCS:
        obj = os.path.join(objdir, s.replace("/", "_").replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append(cu + ["-c", os.path.join(REF, "kernels", s), "-o", obj])
    bind_obj = os.path.join(objdir, "ref_cuda_bindings.o")
    jobs.append(["g++", "-std=c++17", "-O2", "-fPIC", "-w", *defs, *inc, "-c", bind_src, "-o", bind_obj])
    ar_obj = os.path.join(objdir, AR_SRC.replace("/", "_").replace(".cu", ".o"))
    if force or not os.path.exists(ar_obj):
        jobs.append(cu + ["-c", os.path.join(REF, "kernels", AR_SRC), "-o", ar_obj])
    ar_bind_obj = os.path.join(objdir, "ref_cuda_ar_bindings.o")
    ar_defs = [d.replace("_ref_cuda_C", "_ref_cuda_ar_C") for d in defs]
    jobs.append(["g++", "-std=c++17", "-O2", "-fPIC", "-w", *ar_defs, *inc, "-c", ar_bind_src, "-o", ar_bind_obj])
    with ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2))) as ex:
        for cmd, dt in zip(jobs, ex.map(_run, jobs)):
            print(f"  built {os.path.basename(cmd[-1])} in {dt:.0f}s", flush=True)
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", *objs, bind_obj, f"-L{tdir}/lib",
          "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
          "-Xlinker", f"-rpath={tdir}/lib", "-o", TARGET])
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", ar_obj, ar_bind_obj, f"-L{tdir}/lib",
          "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
          "-L/usr/local/cuda/lib64/stubs", "-lcuda", "-Xlinker", f"-rpath={tdir}/lib", "-o", AR_TARGET])
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref/_ref_cuda_C.so:", "ready" if ok else "unavailable (no /root/reference and no prebuilt .so)")
