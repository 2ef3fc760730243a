"""Build libpyticles_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pyticles_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "sph_kernels.cu"), os.path.join(HERE, "csrc", "sph_tiles.cu"),
       os.path.join(HERE, "csrc", "sph_tiles_mma.cu"), os.path.join(HERE, "csrc", "sph_viscous.cu")]
HDR = [os.path.join(ROOT, "include", "pyticles_b200.h"), os.path.join(HERE, "csrc", "sph_device.cuh"),
       os.path.join(HERE, "csrc", "sph_tiles.cuh")]
OUT = os.path.join(HERE, "libpyticles_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include")]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.exists(p) and os.path.getmtime(p) > t for p in SRC + HDR)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + SRC + ["-o", OUT]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(OUT)
