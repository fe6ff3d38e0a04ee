"""In-tree build of the two native artefacts (sm_100a only):

  aphrodite_engine_b200/libb200decode.so   C-ABI CUDA library (nvcc; no torch dependency)
  aphrodite_engine_b200/_C.abi3.so         torch-op shim: TORCH_LIBRARY(_C / _C_cache_ops / _C_cuda_utils)
                                           with the reference's schemas, forwarding to the C ABI

Both are git-ignored and travel to the GPU box with the gpurun snapshot. Rebuilds are incremental
(per-object mtime check against the source and the shared headers).
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build", "obj")
LIB = os.path.join(PKG, "libb200decode.so")
SHIM = os.path.join(PKG, "_C.abi3.so")
CORE = os.path.join(PKG, "_core_C.abi3.so")
MOE = os.path.join(PKG, "_moe_C.abi3.so")

CU_SOURCES = ["runtime.cu", "paged_attention.cu", "cache_ops.cu", "norm_rope_act.cu", "marlin_repack.cu",
              "marlin_gemm.cu", "marlin_gemm_small.cu", "moe_ops.cu", "custom_all_reduce.cu", "misc_ops.cu", "tp_fused.cu",
              "fp8_quant.cu", "scaled_mm.cu", "sampling.cu", "prefill_attention.cu"]
HEADERS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [
    os.path.join(ROOT, "include", "b200_decode.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd, log=None):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if log is not None:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr + "\n")
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    return r


def build_lib(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    jobs, objs = [], []
    for s in CU_SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if _stale(obj, [src] + HEADERS):
            jobs.append(([NVCC, *NVCC_FLAGS, "-c", src, "-o", obj], obj + ".log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    if jobs or _stale(LIB, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
              "-lcudart"])
    if verbose:
        print("built", LIB)
    return LIB


def build_shim(verbose=False, name="_C"):
    import torch

    srcs = [os.path.join(CSRC, f) for f in (("torch_shim.cpp", "torch_shim_r2.cpp") if name == "_C" else ("moe_shim.cpp",))]
    SHIM = os.path.join(PKG, f"{name}.abi3.so")
    if not _stale(SHIM, srcs + [LIB] + HEADERS):
        return SHIM
    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    pyinc = sysconfig.get_paths()["include"]
    cmd = [
        "g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
        "-DTORCH_EXTENSION_NAME=_C", f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include",
        f"-I{pyinc}", "-I/usr/local/cuda/include", f"-I{ROOT}/include", *srcs, "-o", SHIM,
        f"-L{PKG}", "-lb200decode", f"-L{tdir}/lib", "-ltorch", "-ltorch_cpu", "-ltorch_cuda",
        "-lc10", "-lc10_cuda", "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart",
        "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tdir}/lib",
    ]
    _run(cmd, os.path.join(OBJ, f"shim{name}.log"))
    if verbose:
        print("built", SHIM)
    return SHIM


def build_core(verbose=False):
    """`_core_C.abi3.so`: this repo's own ScalarType custom class (csrc/core_scalar_type.cpp)."""
    import torch

    src = os.path.join(CSRC, "core_scalar_type.cpp")
    if not _stale(CORE, [src]):
        return CORE
    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    os.makedirs(OBJ, exist_ok=True)
    _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
          f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include",
          f"-I{sysconfig.get_paths()['include']}", src, "-o", CORE, f"-L{tdir}/lib", "-ltorch",
          "-ltorch_cpu", "-lc10", "-ltorch_python", f"-Wl,-rpath,{tdir}/lib"],
         os.path.join(OBJ, "core_scalar_type.log"))
    if verbose:
        print("built", CORE)
    return CORE


def build(verbose=False):
    build_lib(verbose)
    build_core(verbose)
    build_shim(verbose, "_C")
    build_shim(verbose, "_moe_C")


if __name__ == "__main__":
    build(verbose=True)
