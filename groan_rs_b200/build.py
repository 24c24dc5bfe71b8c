"""Builds libgroan_gpu.so (hand-written CUDA for sm_100a + the C ABI of include/groan_gpu.h) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this also runs in the GPU-less build container.  The library is built
with -fmad=false: the reference (rustc) never contracts a*b+c, and the parity-critical per-atom arithmetic
must round like it; FMAs wanted for speed are written explicitly in the kernels.

Every csrc/*.cu is one translation unit; they are compiled in parallel into build/obj/ and linked (relocatable
device code: the single-pass kernels tail-launch their fallback passes from the device).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libgroan_gpu.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-Xcompiler", "-fPIC,-pthread", "-rdc=true"]
LDFLAGS = ARCH + ["-shared", "-Xcompiler", "-fPIC", "-rdc=true", "-lcudadevrt"]


def units():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    h = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h", ".inl")))
    return h + [os.path.join(ROOT, "include", "groan_gpu.h")]


def _obj(cu):
    return os.path.join(OBJ, os.path.basename(cu)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB, units() + headers())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = headers()
    todo = [cu for cu in units() if force or _stale(_obj(cu), [cu] + hdrs)]

    def compile_one(cu):
        cmd = [nvcc] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", cu, "-o", _obj(cu)]
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, todo))
    subprocess.check_call([nvcc] + LDFLAGS + ["-o", LIB] + [_obj(cu) for cu in units()])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
