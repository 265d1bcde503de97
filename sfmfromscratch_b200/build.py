"""Build libsfmb200.so in-tree with nvcc for sm_100a (B200) only."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libsfmb200.so")
SOURCES = ["api.cu", "extract.cu", "match.cu", "match_tc.cu", "ingest.cu", "ransac.cu", "assoc.cu"]
HEADERS = ["common.cuh", "extract.cuh", "harris_stream.cuh", "tma.cuh", "match.cuh", "ransac_math.cuh", os.path.join("..", "..", "include", "sfmb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-cudart", "static"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsfmb200.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source and link the shared library (incremental)."""
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objdir = os.path.join(_HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.run(cmd, check=True)
        objs.append(o)
    if force or _stale(LIB, objs):
        subprocess.run([nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", LIB] + objs, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
