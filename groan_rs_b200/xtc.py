"""xtc frames in and out of the GPU path (include/groan_xtc.h): the host side of SURVEY.md 8f ranks 2 and 4.

`XtcFile` holds an xtc file's bytes (optionally in pinned memory) and the offsets of its frames; batches of frames are
decoded by a pool of host threads (`decode`) or handed to the GPU as they lie in the file (`System.set_frames_xtc`).
`encode` writes frames exactly like the reference's XtcWriter (src/io/xtc_io/mod.rs:300-330 -> write_xtc).
All work happens in libgroan_gpu.so; this module only marshals arguments.
"""
import ctypes as C
import os

import numpy as np

from . import _lib

_ERR = {1: "end of data", 2: "not an xtc frame (bad magic number)", 3: "data end inside a frame", 4: "damaged frame",
        5: "frames of <= 9 atoms are stored as plain floats (no integer lattice)", 6: "output buffer too small",
        7: "frame does not fit 16-bit lattice offsets", 8: "invalid argument"}


class XtcError(Exception):
    def __init__(self, status, what):
        super().__init__("%s: %s" % (what, _ERR.get(status, "status %d" % status)))
        self.status = status


def _chk(st, what):
    if st != 0:
        raise XtcError(st, what)


def _ptr(a):
    if a is None:
        return None
    if type(a).__module__.startswith("torch"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def default_threads():
    return max(1, min(os.cpu_count() or 1, 32))


class XtcFile:
    """An xtc trajectory held in memory.  `data`: bytes, a numpy uint8 array or a (pinned) torch uint8 tensor."""

    def __init__(self, data, max_frames=1 << 24):
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        self.data = data
        self.nbytes = int(data.numel()) if hasattr(data, "numel") else int(data.size)
        lib = _lib.lib()
        # two passes would need the count first: grow the offsets array instead
        cap = 1024
        while True:
            offsets = np.zeros(cap + 1, np.uint64)
            natoms, n = C.c_int32(0), C.c_size_t(0)
            _chk(lib.groan_xtc_scan(_ptr(data), self.nbytes, min(cap, max_frames), _ptr(offsets), C.byref(natoms), C.byref(n)), "xtc scan")
            if n.value < cap or cap >= max_frames:
                break
            cap *= 8
        self.n_frames = int(n.value)
        self.offsets = offsets[: self.n_frames + 1].copy()
        self.n_atoms = int(natoms.value)

    @classmethod
    def open(cls, path, pinned=False):
        """read a file into memory; pinned=True puts the bytes into page-locked memory (torch), so that
        System.set_frames_xtc uploads them with one asynchronous copy"""
        raw = np.fromfile(path, dtype=np.uint8)
        if pinned:
            import torch
            t = torch.empty(raw.size, dtype=torch.uint8).pin_memory()
            t.numpy()[:] = raw
            return cls(t)
        return cls(raw)

    def decode(self, first=0, count=None, atoms=None, want="xyz", n_threads=None, out=None, origin_out=None):
        """Decode frames [first, first + count) with a pool of host threads.

        want: "xyz" -> floats exactly as read_xtc returns them; "q32" -> the integer lattice; "q16" -> int16 lattice relative
        to a per-frame origin (what System.set_frames_quantized uploads).  atoms: ascending indices = partial-frame read
        (GroupXtcReader).  Returns a dict with the coordinates under `want` plus box [F,9], step, time, precision
        (and origin [F,3] for q16)."""
        count = self.n_frames - first if count is None else count
        if first < 0 or count < 0 or first + count > self.n_frames:
            raise IndexError("frames [%d, %d) of %d" % (first, first + count, self.n_frames))
        sel = None if atoms is None else np.ascontiguousarray(atoms, dtype=np.uint32)
        n_out = self.n_atoms if sel is None else int(sel.size)
        dt = {"xyz": np.float32, "q32": np.int32, "q16": np.int16}[want]
        if out is None:
            out = np.empty((count, n_out, 3), dt)
        res = {want: out, "box": np.zeros((count, 9), np.float32), "step": np.zeros(count, np.int32),
               "time": np.zeros(count, np.float32), "precision": np.zeros(count, np.float32)}
        origin = None
        if want == "q16":
            origin = origin_out if origin_out is not None else np.zeros((count, 3), np.int32)
            res["origin"] = origin
        offs = np.ascontiguousarray(self.offsets[first:first + count + 1])
        _chk(_lib.lib().groan_xtc_decode(_ptr(self.data), self.nbytes, _ptr(offs), count, int(n_threads or default_threads()),
                                         _ptr(sel), 0 if sel is None else int(sel.size),
                                         _ptr(out) if want == "xyz" else None, _ptr(out) if want == "q32" else None,
                                         _ptr(out) if want == "q16" else None, _ptr(origin), _ptr(res["box"]), _ptr(res["step"]),
                                         _ptr(res["time"]), _ptr(res["precision"])), "xtc decode")
        return res


def encode(xyz=None, q=None, boxes=None, step=None, time=None, precision=1000.0, n_threads=None):
    """Frames -> xtc bytes (numpy uint8), byte-identical to the reference's writer.  xyz [F,N,3] floats (quantised like
    write_xtc does) or q [F,N,3] int32 lattice points; boxes [F,9] row-major matrices (or [F,3] / [3] lengths)."""
    from .system import _boxes_to_matrices
    a = np.ascontiguousarray(xyz, dtype=np.float32) if xyz is not None else np.ascontiguousarray(q, dtype=np.int32)
    if a.ndim == 2:
        a = a[None]
    F, N = int(a.shape[0]), int(a.shape[1])
    bm = _boxes_to_matrices(boxes, F)
    st = None if step is None else np.ascontiguousarray(step, dtype=np.int32)
    tm = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
    cap = F * (N * 12 + 128) + 64
    out = np.empty(cap, np.uint8)
    n = C.c_size_t(0)
    _chk(_lib.lib().groan_xtc_encode(_ptr(a) if xyz is not None else None, _ptr(a) if xyz is None else None, F, N, _ptr(bm), _ptr(st),
                                     _ptr(tm), C.c_float(precision), int(n_threads or default_threads()), _ptr(out), cap, C.byref(n)),
         "xtc encode")
    return out[: n.value]
