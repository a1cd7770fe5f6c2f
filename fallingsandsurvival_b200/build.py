"""Builds fallingsandsurvival_b200/libfss_b200.so (the C ABI of include/fss.h) with nvcc for sm_100a.

In-tree, explicit nvcc: the .so travels with the repo snapshot to the GPU box.  `--fmad=false` keeps the
fp32 expressions of the SOUP rule and the temperature stencil un-contracted like the reference's x86-64
-O2 build (no FMA), which bit-exact parity depends on.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfss_b200.so")
SOURCES = [os.path.join(CSRC, "fss_abi.cu")]
DEPS = SOURCES + [os.path.join(CSRC, f) for f in ("fss_device.cuh", "fss_chunkio.h", "fss_rules.cuh", "fss_rules_liq.cuh", "fss_kernels.cuh", "fss_particles.cuh", "fss_render.cuh")] + [
    os.path.join(HERE, "..", "include", f) for f in ("fss.h", "fss_rng.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall", "-shared", "-cudart", "static",
]


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False, extra=()):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + SOURCES + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else [])
