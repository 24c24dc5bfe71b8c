"""Builds libgroan_gpu.so (hand-written CUDA for sm_100a + the C ABI of include/groan_gpu.h) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this also runs in the GPU-less build container.  The library is built
with -fmad=false: the reference (rustc) never contracts a*b+c, and the parity-critical per-atom arithmetic
must round like it; FMAs wanted for speed are written explicitly in the kernels.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgroan_gpu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
    # relocatable device code + device runtime: the single-pass kernels tail-launch their fallback passes from the device
    "-rdc=true", "-lcudadevrt",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(os.path.dirname(HERE), "include", "groan_gpu.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("GROAN_NVCC_EXTRA", "").split()  # experiments only (e.g. -DGROAN_EXP_NOMATH, profiles/exp/README.md)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "groan_gpu.cu")]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
