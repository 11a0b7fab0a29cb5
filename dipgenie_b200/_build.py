"""In-tree build recipes (nvcc for sm_100a, g++ for the host glue). No JIT cache: the built
shared objects live next to the package so that they travel to the GPU box with the repo snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")

CUDA_LIB = os.path.join(PKG, "libdipgenie_cuda.so")
HOST_LIB = os.path.join(PKG, "libdipgenie_host.so")
CLI_BIN = os.path.join(PKG, "dipgenie")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unknown-pragmas", "--expt-relaxed-constexpr", "--expt-extended-lambda",
]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _sources(sub: str, exts):
    d = os.path.join(CSRC, sub)
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return p


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources("cuda", (".cu", ".cpp"))
    deps = srcs + _sources("cuda", (".h", ".cuh")) + _sources("common", (".h",)) + [os.path.join(ROOT, "include", "dipgenie_cuda.h")]
    if force or _newer(CUDA_LIB, deps):
        cmd = [nvcc_path(), *NVCC_FLAGS, "-shared", "-o", CUDA_LIB, *srcs, "-lgomp"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return CUDA_LIB


def build_host(force: bool = False) -> str:
    srcs = [s for s in _sources("host", (".cpp",)) if not s.endswith("main.cpp")]
    if not srcs:
        return ""
    deps = srcs + _sources("host", (".h",)) + _sources("common", (".h",))
    if force or _newer(HOST_LIB, deps):
        # flags mirror the reference build (Makefile:1-2) because the classifier fit is double
        # arithmetic whose contraction behaviour must match; -march as in oracle/Makefile
        cmd = ["g++", "-O3", "-std=c++17", "-fopenmp", "-pthread", "-march=x86-64-v3", "-mtune=generic", "-fPIC",
               "-shared", "-o", HOST_LIB, *srcs, "-lz", "-lm", "-ldl"]
        subprocess.check_call(cmd)
    return HOST_LIB


def build_cli(force: bool = False) -> str:
    main = os.path.join(CSRC, "host", "main.cpp")
    if not os.path.exists(main):
        return ""
    if force or _newer(CLI_BIN, [main, HOST_LIB, CUDA_LIB]):
        cmd = ["g++", "-O3", "-std=c++17", "-fopenmp", "-pthread", "-march=x86-64-v3", "-mtune=generic", "-o", CLI_BIN, main,
               "-L" + PKG, "-ldipgenie_host", "-ldipgenie_cuda", "-Wl,-rpath,$ORIGIN", "-lz", "-lm", "-ldl"]
        subprocess.check_call(cmd)
    return CLI_BIN


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_cli(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", CUDA_LIB)
