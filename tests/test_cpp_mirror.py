"""The C++ host-side mirror (include/groan_gpu.hpp): it must compile against the C ABI without CUDA headers (CPU check)
and its reference-style unit tests (tests/cpp/test_mirror.cpp) must pass on a GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_mirror")


def _compile():
    from groan_rs_b200 import build as b
    lib = b.build()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "test_mirror.cpp"), "-o", EXE, "-L", os.path.dirname(lib),
                           "-lgroan_gpu", "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def test_cpp_mirror_compiles_and_links():
    assert os.path.exists(_compile())


@pytest.mark.gpu
def test_cpp_mirror_reference_style_tests():
    exe = _compile()
    out = subprocess.run([exe, os.path.join(ROOT, "tests", "golden")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
