"""In-tree build of libgorpcuda.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m gorp_b200.build [--force]

The shared library lands next to this file (gorp_b200/libgorpcuda.so): git-ignored, but it travels to the
GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgorpcuda.so")

HOST_SOURCES = [os.path.join(CSRC, "host", f) for f in
                ("common.cpp", "definition.cpp", "automata.cpp", "capture.cpp", "model.cpp", "fused.cpp", "walktables.cpp")]
CUDA_SOURCES = [os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "kernels", "kernels.cu"),
                os.path.join(CSRC, "kernels", "fast.cu"), os.path.join(CSRC, "kernels", "onepass.cu"),
                os.path.join(CSRC, "kernels", "chunkwalk.cu"), os.path.join(CSRC, "kernels", "dfawalk.cu"),
                os.path.join(CSRC, "kernels", "capwalk.cu")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-Wall", "-shared", "-I", os.path.join(ROOT, "include")]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = list(sources)
    for d in (os.path.join(CSRC, "host"), os.path.join(CSRC, "kernels"), os.path.join(ROOT, "include")):
        deps += [os.path.join(d, f) for f in os.listdir(d)]
    return any(os.path.getmtime(s) > t for s in deps)


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build(force=False, verbose=False):
    srcs = CUDA_SOURCES + HOST_SOURCES
    if not force and not _newer(LIB, srcs):
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
