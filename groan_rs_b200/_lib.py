"""ctypes binding of libgroan_gpu.so -- the C ABI declared in include/groan_gpu.h.

This is the Python twin of the Rust `extern "C"` block shown in INTEGRATION.md (which follows the
reference's own xdrfile binding, src/io/xdrfile.rs:27-100).  There is NO fallback: if the CUDA library
is missing or does not load, importing the product fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgroan_gpu.so")

# status codes (include/groan_gpu.h, enum groan_status)
OK, ENOBOX, ENOTORTHO, EEMPTY, ENOPOS, ENOMASS, EGROUPSIZE, EZEROBOX, ENOGROUP, EINVAL, ECUDA, ENOFRAMES, ENOREF, ECAPACITY = range(14)
GROUP_ALL = -1
MAX_GROUPS = 64
FLAG_TRICLINIC = 1
FLAG_EXACT_ONLY = 2
FLAG_NO_TMA = 4
FLAG_HOST_FALLBACK = 16

_vp = C.c_void_p
_sz = C.c_size_t
_int = C.c_int
_u64 = C.c_uint64
_f = C.c_float

# name -> (restype, argtypes); every symbol include/groan_gpu.h declares (tests/test_abi.py checks the two agree)
SIGNATURES = {
    "groan_gpu_create": (_int, [_int, _sz, _sz, C.POINTER(_vp)]),
    "groan_gpu_destroy": (None, [_vp]),
    "groan_gpu_set_flags": (_int, [_vp, C.c_uint]),
    "groan_gpu_set_stream": (_int, [_vp, _vp]),
    "groan_gpu_sync": (_int, [_vp]),
    "groan_gpu_strerror": (C.c_char_p, [_int]),
    "groan_gpu_last_cuda_error": (C.c_char_p, [_vp]),
    "groan_gpu_error_detail": (_int, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "groan_gpu_launch_count": (_u64, [_vp]),
    "groan_gpu_fallback_frames": (_int, [_vp, C.POINTER(_sz)]),
    "groan_gpu_second_pass_frames": (_int, [_vp, C.POINTER(_sz)]),
    "groan_gpu_set_group": (_int, [_vp, _int, _vp, _sz, _vp]),
    "groan_gpu_push_frames": (_int, [_vp, _vp, _vp, _sz]),
    "groan_gpu_push_frames_quantized": (_int, [_vp, _vp, _int, _vp, _f, _vp, _sz]),
    "groan_gpu_get_frames_quantized": (_int, [_vp, _vp, _f]),
    "groan_gpu_attach_frames": (_int, [_vp, _vp, _vp, _sz]),
    "groan_gpu_set_valid": (_int, [_vp, _vp]),
    "groan_gpu_get_frames": (_int, [_vp, _vp]),
    "groan_gpu_estimate_center": (_int, [_vp, _int, _int, _vp]),
    "groan_gpu_get_center": (_int, [_vp, _int, _int, _vp]),
    "groan_gpu_get_center_naive": (_int, [_vp, _int, _vp]),
    "groan_gpu_group_distance": (_int, [_vp, _int, _int, _int, _vp]),
    "groan_gpu_all_distances": (_int, [_vp, _int, _int, _int, _vp]),
    "groan_gpu_all_distances_reduce": (_int, [_vp, _int, _int, _int, _f, _vp, _vp, _vp, _vp, _vp]),
    "groan_gpu_pairs_within": (_int, [_vp, _int, _int, _f, _vp, _vp, _vp, _sz]),
    "groan_gpu_guess_bonds": (_int, [_vp, _vp, _f, _vp, _vp, _sz]),
    "groan_gpu_hbonds": (_int, [_vp, _int, _vp, _vp, _vp, _sz, _f, _f, _vp, _vp, _vp, _sz]),
    "groan_gpu_wrap": (_int, [_vp, _int, _vp]),
    "groan_gpu_translate": (_int, [_vp, _int, C.POINTER(_f), _vp]),
    "groan_gpu_make_group_whole": (_int, [_vp, _int]),
    "groan_gpu_set_molecules": (_int, [_vp, _vp]),
    "groan_gpu_make_molecules_whole": (_int, [_vp]),
    "groan_gpu_atoms_center": (_int, [_vp, _int, _int, _int]),
    "groan_gpu_rmsd_set_reference": (_int, [_vp, _int, _vp, _sz, _vp, _sz, _vp, _vp]),
    "groan_gpu_rmsd": (_int, [_vp, _int, _vp, _vp]),
    "groan_gpu_center_rmsd": (_int, [_vp, _int, _int, _vp, _vp, _vp]),
    "groan_gpu_rmsd_fit": (_int, [_vp, _int, _vp]),
    "groan_gpu_synth_uniform": (_int, [_vp, _u64, _u64, _sz, C.POINTER(_f), C.POINTER(_f), _vp]),
    "groan_gpu_synth_blob": (_int, [_vp, _u64, _u64, _sz, _f, _f, _vp, _vp, _vp, _int]),
    "groan_gpu_synth_blob_ref": (_int, [_vp, _u64, _f, C.POINTER(_f), _vp]),
    # include/groan_xtc.h
    "groan_xtc_scan": (_int, [_vp, _sz, _sz, _vp, C.POINTER(C.c_int32), C.POINTER(_sz)]),
    "groan_xtc_decode": (_int, [_vp, _sz, _vp, _sz, _int, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "groan_xtc_encode": (_int, [_vp, _vp, _sz, _sz, _vp, _vp, _vp, _f, _int, _vp, _sz, C.POINTER(_sz)]),
    "groan_gpu_push_xtc": (_int, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    "groan_gpu_xtc_bad_frames": (_int, [_vp, C.POINTER(_sz)]),
    "groan_gpu_push_group_frames": (_int, [_vp, _vp, _vp, _sz, _vp, _sz]),
    "groan_gpu_write_xtc": (_int, [_vp, _f, _vp, _vp, _int, _vp, _sz, C.POINTER(_sz)]),
}
XTC_OK, XTC_EOF, XTC_EMAGIC, XTC_ETRUNC, XTC_EFORMAT, XTC_ERAW, XTC_ECAPACITY, XTC_ERANGE, XTC_EINVAL = range(9)

_LIB = None


class GroanLibraryMissing(ImportError):
    pass


def lib():
    """Load libgroan_gpu.so (once).  Raises GroanLibraryMissing -- never falls back to a CPU path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise GroanLibraryMissing(
            "%s is missing: build it with `python -m groan_rs_b200.build` (or __graft_entry__.build()); "
            "groan_rs_b200 has no CPU fallback" % LIB_PATH)
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise GroanLibraryMissing("cannot load %s: %s (groan_rs_b200 has no CPU fallback)" % (LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def strerror(status):
    return lib().groan_gpu_strerror(int(status)).decode()
