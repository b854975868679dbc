"""Builds libsrb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The built library is git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SRB200_LIB selects another in-tree build of the same sources (kernel tuning experiments)
LIB = os.environ.get("SRB200_LIB") or os.path.join(HERE, "libsrb200.so")
SOURCES = ["srb_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in os.listdir(root):
            if name.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, name)))
    return m


def build(force=False, verbose=False):
    """Compile if the library is missing or older than any source.  Returns the library path."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("SRB200_NVCC_FLAGS", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    try:
        subprocess.check_call(cmd, cwd=CSRC)
    except FileNotFoundError:
        # try the default CUDA location
        cmd[0] = "/usr/local/cuda/bin/nvcc"
        subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
