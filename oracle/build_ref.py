#!/usr/bin/env python3
"""Build the reference's OWN CPU kernels into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE — never imported by the product path. Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__ may use what this builds.

Compiles, *where they lie* under /root/reference (no sources are copied into this repo):
    kernels/cpu/{attention,cache,layernorm,activation,pos_encoding}.cpp   (AVX512 + OpenMP)
    kernels/cpu/torch_bindings.cpp                                        (op registration)
with flags following cmake/cpu_extension.cmake:17-19,53-66 of the reference, into

    oracle/_ref/_ref_cpu_C_<isa>.so      isa in {avx512bf16, avx512}

TORCH_EXTENSION_NAME is set to `_ref_cpu_C` so the ops register under
torch.ops._ref_cpu_C.* / torch.ops._ref_cpu_C_cache_ops.* and can live in the same process as
this repo's own `_C` library (two TORCH_LIBRARY(_C) blocks in one process are an error).

torch_bindings.cpp is compiled WITHOUT -mavx512* so that its `#ifdef __AVX512F__` int8/oneDNN
block (kernels/cpu/torch_bindings.cpp:91-110; needs libdnnl, absent) is skipped; utils.cpp
(needs libnuma) is replaced by the one-function stub in oracle/ref_stub.cpp.

The GPU box has no /root/reference: it uses the prebuilt .so files (git-ignored, but shipped
by gpurun). This script is a no-op there.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("APHRODITE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

ISAS = {
    "avx512bf16": ["-mavx512f", "-mavx512vl", "-mavx512bw", "-mavx512dq", "-mavx512bf16"],
    "avx512": ["-mavx512f", "-mavx512vl", "-mavx512bw", "-mavx512dq"],
}
KERNEL_SRCS = ["attention", "cache", "layernorm", "activation", "pos_encoding"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("reference CPU build failed")


def build(force: bool = False) -> bool:
    """Returns True if oracle/_ref holds a usable build afterwards."""
    targets = {isa: os.path.join(OUT, f"_ref_cpu_C_{isa}.so") for isa in ISAS}
    if not force and all(os.path.exists(t) for t in targets.values()):
        return True
    if not os.path.isdir(os.path.join(REF, "kernels", "cpu")):
        return any(os.path.exists(t) for t in targets.values())

    import torch

    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    pyinc = sysconfig.get_paths()["include"]
    base = [
        "g++", "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-DAPHRODITE_CPU_EXTENSION",
        "-DTORCH_EXTENSION_NAME=_ref_cpu_C", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
        f"-I{REF}/kernels", f"-I{tdir}/include", f"-I{tdir}/include/torch/csrc/api/include",
        f"-I{pyinc}", "-w",
    ]
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    jobs = []
    for isa, flags in ISAS.items():
        for s in KERNEL_SRCS:
            obj = os.path.join(OUT, "obj", f"{s}_{isa}.o")
            jobs.append(base + flags + ["-c", f"{REF}/kernels/cpu/{s}.cpp", "-o", obj])
    bind_obj = os.path.join(OUT, "obj", "torch_bindings.o")
    stub_obj = os.path.join(OUT, "obj", "ref_stub.o")
    jobs.append(base + ["-c", f"{REF}/kernels/cpu/torch_bindings.cpp", "-o", bind_obj])
    jobs.append(base + ["-c", os.path.join(HERE, "ref_stub.cpp"), "-o", stub_obj])
    with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
        list(ex.map(_run, jobs))
    for isa, tgt in targets.items():
        objs = [os.path.join(OUT, "obj", f"{s}_{isa}.o") for s in KERNEL_SRCS]
        _run(["g++", "-shared", "-fopenmp", *objs, bind_obj, stub_obj, f"-L{tdir}/lib",
              "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python", f"-Wl,-rpath,{tdir}/lib",
              "-o", tgt])
    return True


def build_core_ext(force: bool = False):
    """The reference's `_core_C` (ScalarType custom class, kernels/core/torch_bindings.cpp, plain g++)
    into oracle/_ref/aphrodite_ext/_core_C.abi3.so — only needed to IMPORT the reference's Python
    quantisation utilities when generating golden vectors (tests/golden/make_golden_marlin.py)."""
    tgt = os.path.join(OUT, "aphrodite_ext", "_core_C.abi3.so")
    if os.path.exists(tgt) and not force:
        return tgt
    src = os.path.join(REF, "kernels", "core", "torch_bindings.cpp")
    if not os.path.exists(src):
        return tgt if os.path.exists(tgt) else None
    import torch

    tdir = os.path.dirname(torch.__file__)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    os.makedirs(os.path.dirname(tgt), exist_ok=True)
    _run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-DTORCH_EXTENSION_NAME=_core_C",
          f"-D_GLIBCXX_USE_CXX11_ABI={abi}", f"-I{REF}/kernels/core", f"-I{tdir}/include",
          f"-I{tdir}/include/torch/csrc/api/include", f"-I{sysconfig.get_paths()['include']}", src,
          f"-L{tdir}/lib", "-ltorch", "-ltorch_cpu", "-lc10", "-ltorch_python",
          f"-Wl,-rpath,{tdir}/lib", "-o", tgt])
    return tgt


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("reference _core_C:", build_core_ext(force="--force" in sys.argv))
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference and no prebuilt .so)")
