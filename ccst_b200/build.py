"""Build libccst_b200.so in-tree with nvcc for sm_100a.

    python -m ccst_b200.build [--force] [--verbose]

The library is plain CUDA C++ with a C ABI (no torch, no pybind); nvcc
cross-compiles it without a GPU.  Objects go to ccst_b200/csrc/_obj/, the
shared library to ccst_b200/libccst_b200.so (git-ignored; it travels to the
GPU box with the source snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "libccst_b200.so")
SOURCES = ["api.cu", "stats.cu", "layers.cu", "conv_umma_bf16.cu", "conv_umma_f16.cu"]
HEADERS = ["common.cuh", "layers.h", "conv_umma_impl.cuh", "umma_common.cuh", "conv_main.cuh", "conv_smerge.cuh",
           "conv_first.cuh", "conv_ups4.cuh", "conv_last.cuh", "conv_x3.cuh", os.path.join("..", "..", "include", "ccst_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    """dev=True adds -DCCST_DEV: measurement switches (CCST_ABLATE, CCST_PDL) read from the environment.
    The shipped library is built without it and reads nothing from the environment."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-DCCST_DEV"] if dev else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log = os.path.join(OBJ, src + ".log")
        with open(log, "w") as f:
            f.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose:
            print(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv,
                 dev="--dev" in sys.argv)
    print(path)
