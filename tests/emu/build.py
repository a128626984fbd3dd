"""TEST INFRASTRUCTURE ONLY: builds tests/emu/_build/libsgmcmc_emu.so -- the kernel sources of sgmcmc_b200/csrc compiled by g++
against the CPU emulation in tests/emu (include/cuda_runtime.h, emu.cpp), exporting the same C ABI as the product library.

Why it exists: this container has no GPU.  The emulated build lets the GPU parity tests' logic (same cases, same oracle, same
C-ABI calls) run here, inside `-m "not gpu"`, against the very sources nvcc compiles for sm_100a: SIMT code, shared memory,
warp shuffles, cp.async, mbarrier, TMA (bulk, tile, tile::gather4, tensor-map validation) and tcgen05 (TMEM, UMMA descriptors,
128-byte swizzle).  It detects deadlocks, out-of-range shared / global accesses, misaligned TMA operands, reads that miss their
wait (asynchronous engines complete as late as the memory model allows) -- the failure classes that are Xid faults or hangs on
a real device.  It is not evidence of GPU parity and is never loaded by the product: sgmcmc_b200/_lib.py loads only the nvcc
build and raises when that is missing; nothing under sgmcmc_b200/ or bench.py refers to this directory.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sgmcmc_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libsgmcmc_emu.so")
CXXFLAGS = ["-std=c++17", "-O2", "-g", "-fPIC", "-DSGM_EMU=1", "-fno-strict-aliasing", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-variable",
            "-Wno-unused-but-set-variable", "-Wno-sign-compare", "-Wno-unused-function", "-Wno-maybe-uninitialized", "-Wno-array-bounds", "-Wno-psabi",
            "-I", os.path.join(HERE, "include")]
# (source, language flag): the .cu files are plain C++ once the emulation headers stand in for the CUDA ones
UNITS = [(os.path.join(CSRC, "api.cu"), ["-x", "c++"]), (os.path.join(CSRC, "softmax_net.cu"), ["-x", "c++"]), (os.path.join(HERE, "emu.cpp"), []),
         (os.path.join(HERE, "selftest.cpp"), [])]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "sgmcmc_b200.h"), os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "selftest.cpp")]
    inc = os.path.join(HERE, "include")
    return d + [os.path.join(inc, f) for f in os.listdir(inc)]


def lib_path(variant: str = "") -> str:
    return os.path.join(OUT_DIR + ("_" + variant if variant else ""), "libsgmcmc_emu.so")


def build(force: bool = False, variant: str = "", defs=()) -> str:
    """variant / defs: a second build of the same sources with extra -D flags (e.g. the compiled-out kernel generations)."""
    out_dir = os.path.dirname(lib_path(variant))
    lib = lib_path(variant)
    os.makedirs(out_dir, exist_ok=True)
    if not force and os.path.exists(lib) and all(os.path.getmtime(p) <= os.path.getmtime(lib) for p in _deps()):
        return lib
    cxx = os.environ.get("CXX", "g++")

    def one(unit):
        src, lang = unit
        obj = os.path.join(out_dir, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        subprocess.check_call([cxx, *CXXFLAGS, *defs, "-c", *lang, src, "-o", obj], cwd=ROOT)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(one, UNITS))
    # -Bsymbolic: the emulated cudaMalloc / cudaFree / ... must never be interposed by a real libcudart loaded in the same process
    subprocess.check_call([cxx, "-shared", "-Wl,-Bsymbolic", "-o", lib, *objs, "-ldl"], cwd=ROOT)
    return lib


def build_fake_nccl(force: bool = False) -> str:
    """tests/emu/fake_nccl.c -> _build/libnccl_fake.so: the NCCL entry points over POSIX shared memory (multi-process emulated runs)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    src, lib = os.path.join(HERE, "fake_nccl.c"), os.path.join(OUT_DIR, "libnccl_fake.so")
    if force or not os.path.exists(lib) or os.path.getmtime(src) > os.path.getmtime(lib):
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-o", lib, src, "-lrt", "-lpthread"], cwd=ROOT)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
