"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what include/groan_gpu.h
declares, the ctypes binding covers every symbol, and the product never routes through oracle/ or a CPU fallback.
No compute calls (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "groan_gpu.h")
HEADERS = [HEADER, os.path.join(ROOT, "include", "groan_xtc.h")]


def declared_functions():
    names = set()
    for h in HEADERS:
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(groan_(?:gpu|xtc)_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


@pytest.fixture(scope="module")
def libpath():
    from groan_rs_b200 import build
    return build.build()


def test_header_declares_the_survey_boundary():
    names = declared_functions()
    for want in ["groan_gpu_create", "groan_gpu_destroy", "groan_gpu_set_group", "groan_gpu_push_frames",
                 "groan_gpu_estimate_center", "groan_gpu_get_center", "groan_gpu_group_distance", "groan_gpu_all_distances",
                 "groan_gpu_all_distances_reduce", "groan_gpu_wrap", "groan_gpu_translate", "groan_gpu_rmsd_set_reference",
                 "groan_gpu_rmsd", "groan_gpu_rmsd_fit", "groan_gpu_sync", "groan_gpu_strerror"]:
        assert want in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in declared_functions():
        assert hasattr(lib, name), "libgroan_gpu.so does not export %s" % name


def test_binding_covers_every_declared_symbol():
    from groan_rs_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_strerror_needs_no_gpu(libpath):
    from groan_rs_b200 import _lib
    assert _lib.strerror(0) == "ok"
    assert "orthogonal" in _lib.strerror(_lib.ENOTORTHO)
    assert _lib.strerror(12345) == "unknown status"


def test_status_codes_match_header():
    from groan_rs_b200 import _lib
    src = open(HEADER).read()
    for name in ["OK", "ENOBOX", "ENOTORTHO", "EEMPTY", "ENOPOS", "ENOMASS", "EGROUPSIZE", "EZEROBOX", "ENOGROUP", "EINVAL",
                 "ECUDA", "ENOFRAMES", "ENOREF", "ECAPACITY"]:
        m = re.search(r"GROAN_%s\s*=\s*(\d+)" % name, src)
        assert m and int(m.group(1)) == getattr(_lib, name), name


def test_product_is_kernels_only_no_oracle_no_fallback(libpath):
    """The product path must not import, link or call anything under oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "groan_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "libgroan_oracle" not in text and "groan_oracle.h" not in text, f
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    assert "orc_" not in out
    # the library must carry sm_100a device code
    sass = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from groan_rs_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GroanLibraryMissing):
        _lib.lib()


def test_async_copies_keep_their_cache_policy_descriptor(libpath):
    """ptxas 12.9 can emit LDGSTS [R + UR + imm], desc[UR] with the warp-uniform stage offset copied over the cache-policy
    descriptor (B200: 'illegal instruction', profiles/r2_async.md).  stream_quads_warp carries the destination as a per-thread
    address to avoid that form: no cp.async of the RMSD kernels may use a uniform register in its shared-memory address,
    and the kernels must still be fed by LDGSTS with L2 cache hints (desc[...])."""
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "k_rmsd_quad", libpath], capture_output=True, text=True).stdout
    if "LDGSTS" not in sass:  # older cuobjdump: no -fun filter on mangled substrings
        sass = subprocess.run(["cuobjdump", "-sass", libpath], capture_output=True, text=True).stdout
    copies = [l for l in sass.splitlines() if "LDGSTS" in l]
    assert len(copies) >= 14, "the RMSD kernels are fed by cp.async (LDGSTS)"
    assert all("desc[" in l for l in copies), "every async copy carries an L2 cache-policy descriptor"
    bad = [l.strip() for l in copies if re.search(r"\[R\d+\+UR\d+", l)]
    assert not bad, "LDGSTS with a uniform-register shared-memory offset:\n" + "\n".join(bad[:4])
