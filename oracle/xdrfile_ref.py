"""ctypes binding of oracle/_ref/libxdrfile.so -- the REFERENCE's own vendored C xtc codec (external/xdrfile), compiled
from where it lies under /root/reference by `make -C oracle ref`.  TEST INFRASTRUCTURE: only tests/, bench.py's CPU leg and
__graft_entry__.smoke() may use it; it is the checker of groan_rs_b200's own xtc codec (csrc/xtc_codec.hpp, kernels_xtc.cuh).
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libxdrfile.so")
_LIB = None


def available():
    return os.path.exists(PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(PATH)
        L.xdrfile_open.restype = ctypes.c_void_p
        L.xdrfile_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.xdrfile_close.argtypes = [ctypes.c_void_p]
        L.read_xtc_natoms.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.read_xtc.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float),
                               ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        L.write_xtc.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_float]
        _LIB = L
    return _LIB


def read_xtc(path):
    """every frame of an xtc file through the reference's read_xtc (xdrfile_xtc.h:45-52):
    dict(xyz [F,N,3], box [F,9], step, time, precision)"""
    L = lib()
    n = ctypes.c_int(0)
    assert L.read_xtc_natoms(path.encode(), ctypes.byref(n)) == 0, path
    xd = L.xdrfile_open(path.encode(), b"r")
    assert xd
    xyz, box, step, time, prec = [], [], [], [], []
    while True:
        x = np.zeros((n.value, 3), np.float32)
        b = np.zeros((3, 3), np.float32)
        s, t, p = ctypes.c_int(0), ctypes.c_float(0), ctypes.c_float(0)
        rc = L.read_xtc(xd, n.value, ctypes.byref(s), ctypes.byref(t), b.ctypes.data, x.ctypes.data, ctypes.byref(p))
        if rc != 0:
            assert rc == 11, rc  # exdrENDOFFILE
            break
        xyz.append(x), box.append(b.reshape(9)), step.append(s.value), time.append(t.value), prec.append(p.value)
    L.xdrfile_close(xd)
    return {"xyz": np.stack(xyz), "box": np.stack(box), "step": np.array(step, np.int32), "time": np.array(time, np.float32),
            "precision": np.array(prec, np.float32)}


def write_xtc(path, xyz, box, step, time, precision):
    """frames through the reference's write_xtc (xdrfile_xtc.h:55-59); returns the file's bytes"""
    L = lib()
    xd = L.xdrfile_open(path.encode(), b"w")
    assert xd
    xyz = np.ascontiguousarray(xyz, np.float32)
    box = np.ascontiguousarray(box, np.float32).reshape(-1, 9)
    for f in range(xyz.shape[0]):
        x, b = xyz[f].copy(), box[f].copy()
        assert L.write_xtc(xd, xyz.shape[1], int(step[f]), float(time[f]), b.ctypes.data, x.ctypes.data, float(precision)) == 0
    L.xdrfile_close(xd)
    return np.fromfile(path, dtype=np.uint8)
