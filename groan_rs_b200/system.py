"""Host-side mirror of groan_rs's System / Group / SimBox / Dimension surface for the PBC geometry hot path.

Same method names, argument meaning and error behaviour as the reference (citations per method), but
every analysis method evaluates ALL frames of the current batch on the GPU and returns an array with a
leading frame axis.  Everything here is plumbing around the C ABI (include/groan_gpu.h): argument
marshalling, group bookkeeping, error translation.  No arithmetic of the hot path happens in Python and
there is no CPU fallback.
"""
import ctypes as C
import enum
import weakref

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------------------------- errors
class GroanError(Exception):
    """Base of the mirrored error enums; `variant` names the reference's enum variant."""

    def __init__(self, variant, message, status=None, detail=None):
        super().__init__("%s: %s" % (variant, message))
        self.variant = variant
        self.status = status
        self.detail = detail


class GroupError(GroanError):  # src/errors.rs GroupError
    pass


class SimBoxError(GroanError):  # src/errors.rs:556-582
    pass


class PositionError(GroanError):  # src/errors.rs:227-258
    pass


class MassError(GroanError):  # src/errors.rs:290-305
    pass


class RMSDError(GroanError):  # src/errors.rs:624-650
    pass


class GpuError(GroanError):
    pass


class Dimension(enum.IntEnum):
    """src/structures/dimension.rs:15-25"""
    None_ = 0
    X = 1
    Y = 2
    Z = 3
    XY = 4
    XZ = 5
    YZ = 6
    XYZ = 7

    @classmethod
    def from_bools(cls, x, y, z):  # dimension.rs From<[bool;3]>
        return {(0, 0, 0): cls.None_, (1, 0, 0): cls.X, (0, 1, 0): cls.Y, (0, 0, 1): cls.Z, (1, 1, 0): cls.XY,
                (1, 0, 1): cls.XZ, (0, 1, 1): cls.YZ, (1, 1, 1): cls.XYZ}[(int(bool(x)), int(bool(y)), int(bool(z)))]


class SimBox:
    """src/structures/simbox.rs:13-52.  Nine floats in GROMACS order v1x v2y v3z v1y v1z v2x v2z v3x v3y."""

    def __init__(self, values):
        v = np.asarray(values, dtype=np.float32).reshape(-1)
        if v.size == 3:
            v = np.concatenate([v, np.zeros(6, np.float32)])
        if v.size != 9:
            raise ValueError("SimBox takes 3 or 9 floats")
        if v[3] != 0 or v[4] != 0 or v[6] != 0:  # simbox.rs:28-52 panics
            raise ValueError("SimBox: v1y, v1z and v2z must be zero")
        self.v = v.astype(np.float32)

    @classmethod
    def from_matrix(cls, m):
        """io/xdrfile.rs:170-187: row-major 3x3, box[i][j] = component j of box vector i"""
        m = np.asarray(m, dtype=np.float32).reshape(3, 3)
        return cls([m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[0, 2], m[1, 0], m[1, 2], m[2, 0], m[2, 1]])

    def matrix(self):
        v = self.v
        return np.array([v[0], v[3], v[4], v[5], v[1], v[6], v[7], v[8], v[2]], dtype=np.float32)

    @property
    def x(self):
        return self.v[0]

    @property
    def y(self):
        return self.v[1]

    @property
    def z(self):
        return self.v[2]

    def is_orthogonal(self):  # simbox.rs:185
        return self.v[5] == 0 and self.v[7] == 0 and self.v[8] == 0


def _boxes_to_matrices(boxes, n_frames):
    """Accepts None, a SimBox, [3], [9] (row-major matrix), [3,3], [F,3], [F,9], [F,3,3]; returns F x 9 f32 or None."""
    if boxes is None:
        return None
    if isinstance(boxes, SimBox):
        return np.ascontiguousarray(np.tile(boxes.matrix(), (n_frames, 1)), dtype=np.float32)
    b = np.asarray(boxes, dtype=np.float32)
    if b.ndim == 1 and b.size == 3:
        b = np.tile(b, (n_frames, 1))
    elif b.ndim == 1 and b.size == 9:
        b = np.tile(b, (n_frames, 1))
    elif b.ndim == 2 and b.shape == (3, 3) and n_frames != 3:
        b = np.tile(b.reshape(1, 9), (n_frames, 1))
    if b.ndim == 3:
        b = b.reshape(b.shape[0], 9)
    if b.shape[0] != n_frames:
        raise ValueError("boxes: expected %d frames, got %r" % (n_frames, b.shape))
    if b.shape[1] == 3:
        m = np.zeros((n_frames, 9), np.float32)
        m[:, 0], m[:, 4], m[:, 8] = b[:, 0], b[:, 1], b[:, 2]
        b = m
    return np.ascontiguousarray(b, dtype=np.float32)


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _ptr(x):
    if x is None:
        return None
    if _is_torch(x):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(x.ctypes.data)


class Group:
    """src/structures/group.rs, container.rs:51-115: sorted, de-duplicated atom indices; iteration ascending."""

    def __init__(self, indices, n_atoms):
        idx = np.unique(np.asarray(indices, dtype=np.int64).reshape(-1))
        if idx.size and (idx[0] < 0 or idx[-1] >= n_atoms):
            raise ValueError("group index out of range")
        self.indices = np.ascontiguousarray(idx, dtype=np.uint32)
        self.gid = None

    def __len__(self):
        return int(self.indices.size)

    def index_array(self, n_atoms):
        return np.arange(n_atoms, dtype=np.uint32) if self.indices is None else self.indices


class System:
    """A molecular system whose per-frame PBC geometry runs on one B200 (src/system/mod.rs:38-73).

    Frames are given in batches (`set_frames`, the FrameData::update_system tap of xdrfile_xtc.rs:88-104);
    every analysis method returns one result per frame.
    """

    def __init__(self, n_atoms, masses=None, device=0, max_frames=1, triclinic=False):
        self.n_atoms = int(n_atoms)
        self.max_frames = int(max_frames)
        self._lib = _lib.lib()
        h = C.c_void_p()
        self._check(self._lib.groan_gpu_create(int(device), self.n_atoms, self.max_frames, C.byref(h)), "create")
        self._h = h
        self.device = int(device)
        self.masses = None if masses is None else np.ascontiguousarray(masses, dtype=np.float32)
        self._groups = {}
        self._next_gid = 0
        self.n_frames = 0
        self._host_frames = None
        self._boxes = None
        self._version = 0   # bumped by everything that changes positions, boxes or groups (keys the RMSD reference cache)
        self._ref_cache = {}  # target gid -> (weakref to the reference System, its _version when uploaded, group name)
        self._keep = []
        self._all = None
        self._atom_pair = None
        if self.masses is not None:
            # "all" / "All" exist from the start with the atoms' masses (System::new, src/system/mod.rs:150-170)
            gm = np.where(np.isnan(self.masses), np.float32(-1.0), self.masses).astype(np.float32)
            self._check(self._lib.groan_gpu_set_group(self._h, _lib.GROUP_ALL, None, self.n_atoms, _ptr(gm)), "set_group", "all")
        if triclinic:
            self.set_flags(_lib.FLAG_TRICLINIC)

    # ------------------------------------------------------------------ lifetime / errors
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.groan_gpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_flags(self, flags):
        self._check(self._lib.groan_gpu_set_flags(self._h, int(flags)), "set_flags")

    def set_stream(self, cuda_stream):
        """Run all kernels on a caller-owned CUDA stream (an int handle, e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._lib.groan_gpu_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else 0)), "set_stream")

    def sync(self):
        self._check(self._lib.groan_gpu_sync(self._h), "sync")

    def launch_count(self):
        return int(self._lib.groan_gpu_launch_count(self._h))

    def fallback_frames(self):
        """frames of the last centre / RMSD call that the single-pass kernel handed to the reference-order passes"""
        n = C.c_size_t(0)
        self._check(self._lib.groan_gpu_fallback_frames(self._h, C.byref(n)), "fallback_frames")
        return int(n.value)

    def second_pass_frames(self):
        """frames of the last group_center_and_rmsd call whose centre went through the sine-sum pass (second tier)"""
        n = C.c_size_t(0)
        self._check(self._lib.groan_gpu_second_pass_frames(self._h, C.byref(n)), "second_pass_frames")
        return int(n.value)

    def _detail(self):
        a, b = C.c_size_t(0), C.c_size_t(0)
        self._lib.groan_gpu_error_detail(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def _check(self, st, what, group=None, rmsd=False):
        if st == _lib.OK:
            return
        msg = _lib.strerror(st)
        L = _lib
        if st == L.ECUDA:
            raise GpuError("CudaError", self._lib.groan_gpu_last_cuda_error(self._h).decode(), st)
        if st in (L.ENOBOX, L.ENOTORTHO, L.EZEROBOX):
            variant = {L.ENOBOX: "SimBoxError::DoesNotExist", L.ENOTORTHO: "SimBoxError::NotOrthogonal",
                       L.EZEROBOX: "panic: Box len should not be zero"}[st]
            outer = RMSDError if rmsd else (GroupError if group is not None else SimBoxError)
            e = outer("InvalidSimBox(%s)" % variant, msg, st)
            e.inner = SimBoxError(variant, msg, st)
            raise e
        if st == L.EEMPTY:
            raise (RMSDError if rmsd else GroupError)("EmptyGroup", "%s (%s)" % (msg, group), st)
        if st == L.ENOGROUP:
            raise (RMSDError("NonexistentGroup", "%s (%s)" % (msg, group), st) if rmsd
                   else GroupError("NotFound", "%s (%s)" % (msg, group), st))
        if st == L.ENOPOS:
            f, a = self._detail()
            e = (RMSDError if rmsd else GroupError)("InvalidPosition(PositionError::NoPosition(%d))" % a, msg, st, (f, a))
            e.inner = PositionError("NoPosition", "atom %d has no position (frame %d)" % (a, f), st, (f, a))
            raise e
        if st == L.ENOMASS:
            f, a = self._detail()
            e = (RMSDError if rmsd else GroupError)("InvalidMass(MassError::NoMass(%d))" % a, msg, st, (f, a))
            e.inner = MassError("NoMass", "atom %d has no mass" % a, st, (f, a))
            raise e
        if st == L.EGROUPSIZE:
            a, b = self._detail()
            raise RMSDError("InconsistentGroup", "%s: %d atoms in reference, %d in target (%s)" % (msg, a, b, group), st, (a, b))
        raise GpuError({L.EINVAL: "InvalidArgument", L.ENOFRAMES: "NoFrames", L.ENOREF: "NoReference",
                        L.ECAPACITY: "Capacity"}.get(st, "Unknown"), "%s in %s" % (msg, what), st)

    # ------------------------------------------------------------------ groups (src/system/groups.rs)
    def group_create_from_indices(self, name, indices):
        """Group::from_indices (group.rs:94): sorted + de-duplicated.  Re-creating a name overwrites it."""
        g = Group(indices, self.n_atoms)
        if name in self._groups:
            g.gid = self._groups[name].gid
        else:
            if self._next_gid >= _lib.MAX_GROUPS:
                raise GpuError("Capacity", "too many groups")
            g.gid = self._next_gid
            self._next_gid += 1
        gm = None
        if self.masses is not None:
            gm = np.ascontiguousarray(self.masses[g.indices.astype(np.int64)], dtype=np.float32)
            gm = np.where(np.isnan(gm), np.float32(-1.0), gm).astype(np.float32)
        g.mass = gm
        self._check(self._lib.groan_gpu_set_group(self._h, g.gid, _ptr(g.indices) if len(g) else None, len(g),
                                                  _ptr(gm) if gm is not None and len(g) else None), "set_group", name)
        self._groups[name] = g
        self._ref_cache.pop(g.gid, None)  # set_group dropped the device-side reference of this gid
        self._version += 1                # ... and a System used as somebody's reference has changed
        return g

    def group_create_from_ranges(self, name, ranges):
        """container.rs:13-31: inclusive (start, end) ranges"""
        idx = np.concatenate([np.arange(a, b + 1) for a, b in ranges]) if ranges else np.zeros(0, np.int64)
        return self.group_create_from_indices(name, idx)

    def group_exists(self, name):
        return name in ("all", "All") or name in self._groups

    def group_get_n_atoms(self, name):
        return self.n_atoms if name in ("all", "All") and name not in self._groups else len(self._group(name))

    def _bump(self):
        """positions changed in place: whoever cached this System as an RMSD reference must upload it again"""
        self._host_frames = None
        self._version += 1

    def group_isempty(self, name):
        return self.group_get_n_atoms(name) == 0

    def _group(self, name, rmsd=False):
        if name in ("all", "All") and name not in self._groups:
            if self._all is None:
                self._all = Group.__new__(Group)
                self._all.indices, self._all.gid = None, _lib.GROUP_ALL  # indices None: every atom, ascending
                self._all.mass = None if self.masses is None else np.where(np.isnan(self.masses), np.float32(-1.0),
                                                                          self.masses).astype(np.float32)
            return self._all
        if name not in self._groups:
            if rmsd:
                raise RMSDError("NonexistentGroup", name, _lib.ENOGROUP)
            raise GroupError("NotFound", name, _lib.ENOGROUP)
        return self._groups[name]

    def _gid(self, name, rmsd=False):
        if name in ("all", "All") and name not in self._groups:
            return _lib.GROUP_ALL
        return self._group(name, rmsd).gid

    # ------------------------------------------------------------------ frames
    def set_frames(self, xyz, boxes):
        """Stage a batch: xyz [F, N, 3] (or [N, 3]) f32 as read_xtc emits; boxes per frame (see _boxes_to_matrices).

        numpy / pinned torch input is copied host->device on the copy stream (overlapping kernels still running on
        the previous batch); a CUDA torch tensor is attached zero-copy and modified in place by wrap / fit.
        """
        if _is_torch(xyz):
            t = xyz if xyz.dim() == 3 else xyz.unsqueeze(0)
            if str(t.dtype) != "torch.float32" or not t.is_contiguous():
                raise ValueError("frames must be contiguous float32")
            F = int(t.shape[0])
            if tuple(t.shape[1:]) != (self.n_atoms, 3):
                raise ValueError("frames must be [F, %d, 3]" % self.n_atoms)
            bm = _boxes_to_matrices(boxes, F)
            fn = self._lib.groan_gpu_attach_frames if t.is_cuda else self._lib.groan_gpu_push_frames
            self._check(fn(self._h, _ptr(t), _ptr(bm), F), "set_frames")
            self._keep = [t]
            self._host_frames = None
        else:
            a = np.ascontiguousarray(xyz, dtype=np.float32)
            if a.ndim == 2:
                a = a[None]
            F = int(a.shape[0])
            if a.shape[1:] != (self.n_atoms, 3):
                raise ValueError("frames must be [F, %d, 3]" % self.n_atoms)
            bm = _boxes_to_matrices(boxes, F)
            self._check(self._lib.groan_gpu_push_frames(self._h, _ptr(a), _ptr(bm), F), "set_frames")
            self._host_frames = a
            self._keep = [a]
        self.n_frames = F
        self._boxes = bm
        self._version += 1

    def set_frames_quantized(self, q, precision, boxes, origin=None):
        """Stage a batch in the form the xtc decoder holds it before it emits floats: q [F, N, 3] int16 or int32 lattice
        points, optional per-frame integer origin [F, 3]; coordinate = (float)(q + origin) * (1 / precision) exactly as
        external/xdrfile/xdrfile.c:844,915-917 computes it.  Half (int16) the PCIe bytes of set_frames for the same floats."""
        if _is_torch(q):
            t = q if q.dim() == 3 else q.unsqueeze(0)
            eb = {"torch.int16": 2, "torch.int32": 4}[str(t.dtype)]
            if not t.is_contiguous() or t.is_cuda:
                raise ValueError("quantised frames must be contiguous host tensors")
            a, F, shape = t, int(t.shape[0]), tuple(t.shape[1:])
        else:
            a = np.ascontiguousarray(q)
            if a.dtype not in (np.int16, np.int32):
                raise ValueError("quantised frames must be int16 or int32")
            if a.ndim == 2:
                a = a[None]
            eb, F, shape = a.dtype.itemsize, int(a.shape[0]), a.shape[1:]
        if tuple(shape) != (self.n_atoms, 3):
            raise ValueError("frames must be [F, %d, 3]" % self.n_atoms)
        o = None if origin is None else np.ascontiguousarray(origin, dtype=np.int32).reshape(F, 3)
        bm = _boxes_to_matrices(boxes, F)
        self._check(self._lib.groan_gpu_push_frames_quantized(self._h, _ptr(a), eb, _ptr(o), C.c_float(precision), _ptr(bm), F),
                    "set_frames_quantized")
        self._host_frames = None
        self._keep = [a, o]
        self.n_frames = F
        self._boxes = bm
        self._version += 1

    def set_frames_xtc(self, xtc, first=0, count=None):
        """Stage frames [first, first + count) of an xtc.XtcFile straight from the file's bytes: they cross PCIe compressed
        (about a third of the floats for a solvated system) and are decoded on the GPU (groan_gpu_push_xtc); boxes come from
        the frame headers.  Returns dict(step, time, precision)."""
        count = xtc.n_frames - first if count is None else int(count)
        if first < 0 or count <= 0 or first + count > xtc.n_frames:
            raise IndexError("frames [%d, %d) of %d" % (first, first + count, xtc.n_frames))
        offs = np.ascontiguousarray(xtc.offsets[first:first + count + 1])
        meta = {"step": np.zeros(count, np.int32), "time": np.zeros(count, np.float32), "precision": np.zeros(count, np.float32)}
        self._check(self._lib.groan_gpu_push_xtc(self._h, _ptr(xtc.data), xtc.nbytes, _ptr(offs), count, _ptr(meta["step"]),
                                                 _ptr(meta["time"]), _ptr(meta["precision"])), "set_frames_xtc")
        self._host_frames = None
        self._keep = [xtc.data, offs]
        self.n_frames = count
        bm = np.zeros((count, 9), np.float32)
        for f in range(count):  # the header's box: 9 big-endian floats at byte 16 of the frame
            o = int(offs[f]) + 16
            raw = xtc.data[o:o + 36]
            raw = raw.numpy() if _is_torch(raw) else np.asarray(raw)
            bm[f] = np.frombuffer(raw.tobytes(), dtype=">f4")
        self._boxes = bm
        self._version += 1
        return meta

    def xtc_bad_frames(self):
        n = C.c_size_t(0)
        self._check(self._lib.groan_gpu_xtc_bad_frames(self._h, C.byref(n)), "xtc_bad_frames")
        return int(n.value)

    def set_group_frames(self, xyz_sel, atoms, boxes):
        """Partial frames (GroupXtcReader, molly_xtc.rs:404-470): xyz_sel [F, len(atoms), 3] holds only the atoms `atoms`
        (ascending).  Only those bytes cross PCIe; the other atoms of the System keep their previous values."""
        sel = np.ascontiguousarray(atoms, dtype=np.uint32)
        if _is_torch(xyz_sel):
            a = xyz_sel if xyz_sel.dim() == 3 else xyz_sel.unsqueeze(0)
            if str(a.dtype) != "torch.float32" or not a.is_contiguous() or a.is_cuda:
                raise ValueError("partial frames must be contiguous float32 host tensors")
        else:
            a = np.ascontiguousarray(xyz_sel, dtype=np.float32)
            if a.ndim == 2:
                a = a[None]
        F = int(a.shape[0])
        if tuple(a.shape[1:]) != (int(sel.size), 3):
            raise ValueError("partial frames must be [F, %d, 3]" % sel.size)
        bm = _boxes_to_matrices(boxes, F)
        self._check(self._lib.groan_gpu_push_group_frames(self._h, _ptr(a), _ptr(sel), int(sel.size), _ptr(bm), F), "set_group_frames")
        self._host_frames = None
        self._keep = [a, sel]
        self.n_frames = F
        self._boxes = bm
        self._version += 1

    def write_xtc(self, precision=1000.0, step=None, time=None, n_threads=None):
        """The current batch as xtc frames (numpy uint8), byte-identical to XtcWriter::write_frame (xtc_io/mod.rs:300-330):
        quantised on the device with the writer's rounding, encoded by a pool of host threads."""
        from . import xtc as _xtc
        F = self.n_frames
        st = None if step is None else np.ascontiguousarray(step, dtype=np.int32)
        tm = None if time is None else np.ascontiguousarray(time, dtype=np.float32)
        cap = F * (self.n_atoms * 12 + 128) + 64
        out = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        self._check(self._lib.groan_gpu_write_xtc(self._h, C.c_float(precision), _ptr(st), _ptr(tm),
                                                  int(n_threads or _xtc.default_threads()), _ptr(out), cap, C.byref(n)), "write_xtc")
        return out[: n.value]

    def set_valid(self, valid):
        """Option<Vector3D> positions (atom.rs:23-71): valid[f, i] == 0 means atom i has no position in frame f."""
        v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8).reshape(self.n_frames, self.n_atoms)
        self._check(self._lib.groan_gpu_set_valid(self._h, _ptr(v)), "set_valid")

    def get_frames(self, out=None):
        if out is None:
            out = np.empty((self.n_frames, self.n_atoms, 3), np.float32)
        self._check(self._lib.groan_gpu_get_frames(self._h, _ptr(out)), "get_frames")
        return out

    def get_frames_quantized(self, precision, out=None):
        """the batch as the xtc encoder's integer lattice points, (int)(x * precision +- 0.5) in f32 (xdrfile.c:1018-1031)"""
        if out is None:
            out = np.empty((self.n_frames, self.n_atoms, 3), np.int32)
        self._check(self._lib.groan_gpu_get_frames_quantized(self._h, _ptr(out), C.c_float(precision)), "get_frames_quantized")
        return out

    def synth_uniform(self, seed, frame0, n_frames, lo, span, boxes):
        bm = _boxes_to_matrices(boxes, n_frames)
        lo3 = (C.c_float * 3)(*[float(v) for v in lo])
        sp3 = (C.c_float * 3)(*[float(v) for v in span])
        self._check(self._lib.groan_gpu_synth_uniform(self._h, seed, frame0, n_frames, lo3, sp3, _ptr(bm)), "synth_uniform")
        self.n_frames, self._boxes, self._host_frames = n_frames, bm, None
        self._version += 1

    def synth_blob(self, seed, frame0, n_frames, scale, nscale, rot, centre, boxes, wrap=True):
        bm = _boxes_to_matrices(boxes, n_frames)
        r = np.ascontiguousarray(rot, dtype=np.float32).reshape(n_frames, 9)
        c = np.ascontiguousarray(centre, dtype=np.float32).reshape(n_frames, 3)
        self._check(self._lib.groan_gpu_synth_blob(self._h, seed, frame0, n_frames, scale, nscale, _ptr(r), _ptr(c), _ptr(bm),
                                                   1 if wrap else 0), "synth_blob")
        self.n_frames, self._boxes, self._host_frames = n_frames, bm, None
        self._version += 1

    def synth_blob_ref(self, seed, scale, centre, out=None):
        if out is None:
            out = np.empty((self.n_atoms, 3), np.float32)
        c3 = (C.c_float * 3)(*[float(v) for v in centre])
        self._check(self._lib.groan_gpu_synth_blob_ref(self._h, seed, scale, c3, _ptr(out)), "synth_blob_ref")
        return out

    def get_box_center(self):
        """System::get_box_center (mod.rs:298-308), per frame"""
        if self._boxes is None:
            raise SimBoxError("SimBoxError::DoesNotExist", "no box", _lib.ENOBOX)
        b = self._boxes
        if np.any(b[:, [3, 6, 7]] != 0):
            raise SimBoxError("SimBoxError::NotOrthogonal", "box not orthogonal", _lib.ENOTORTHO)
        return np.stack([b[:, 0] / np.float32(2), b[:, 4] / np.float32(2), b[:, 8] / np.float32(2)], axis=1)

    # ------------------------------------------------------------------ outputs
    def _out(self, out, shape, dtype=np.float32):
        if out is None:
            return np.empty(shape, dtype)
        return out

    # ------------------------------------------------------------------ centres (src/system/analysis.rs)
    def group_estimate_center(self, name, out=None):
        """analysis.rs:52 -> iterators.rs:1152-1191 (plain Bai-Breen)"""
        out = self._out(out, (self.n_frames, 3))
        self._check(self._lib.groan_gpu_estimate_center(self._h, self._gid(name), 0, _ptr(out)), "group_estimate_center", name)
        return out

    def group_estimate_com(self, name, out=None):
        """iterators.rs:1314-1357"""
        out = self._out(out, (self.n_frames, 3))
        self._check(self._lib.groan_gpu_estimate_center(self._h, self._gid(name), 1, _ptr(out)), "group_estimate_com", name)
        return out

    def group_get_center(self, name, out=None):
        """analysis.rs:105-120 -> iterators.rs:1237-1266 (refined Bai-Breen)"""
        out = self._out(out, (self.n_frames, 3))
        self._check(self._lib.groan_gpu_get_center(self._h, self._gid(name), 0, _ptr(out)), "group_get_center", name)
        return out

    def group_get_com(self, name, out=None):
        """analysis.rs:258 -> iterators.rs:1404-1438"""
        out = self._out(out, (self.n_frames, 3))
        self._check(self._lib.groan_gpu_get_center(self._h, self._gid(name), 1, _ptr(out)), "group_get_com", name)
        return out

    def group_get_center_naive(self, name, out=None):
        """iterators.rs:886-903"""
        out = self._out(out, (self.n_frames, 3))
        self._check(self._lib.groan_gpu_get_center_naive(self._h, self._gid(name), _ptr(out)), "group_get_center_naive", name)
        return out

    # ------------------------------------------------------------------ distances
    def group_distance(self, group1, group2, dim, out=None):
        """analysis.rs:348-360"""
        g1, g2 = self._gid(group1), self._gid(group2)
        out = self._out(out, (self.n_frames,))
        self._check(self._lib.groan_gpu_group_distance(self._h, g1, g2, int(dim), _ptr(out)), "group_distance",
                    "%s/%s" % (group1, group2))
        return out

    def group_all_distances(self, group1, group2, dim, out=None):
        """analysis.rs:401-427: [F, n1, n2] row-major"""
        g1, g2 = self._gid(group1), self._gid(group2)
        n1, n2 = self.group_get_n_atoms(group1), self.group_get_n_atoms(group2)
        out = self._out(out, (self.n_frames, n1, n2))
        self._check(self._lib.groan_gpu_all_distances(self._h, g1, g2, int(dim), _ptr(out) if n1 * n2 else None),
                    "group_all_distances", "%s/%s" % (group1, group2))
        return out

    def atoms_distance(self, index1, index2, dim):
        """System::atoms_distance (analysis.rs:459-471) -> Atom::distance (atom.rs:780-790), per frame [F]; the first atom's
        position is checked first, like the reference"""
        for i in (index1, index2):
            if not 0 <= int(i) < self.n_atoms:
                raise IndexError("atom index %d out of range" % i)
        if self._atom_pair != (int(index1), int(index2)):
            self.group_create_from_indices("__atom1", [int(index1)])
            self.group_create_from_indices("__atom2", [int(index2)])
            self._atom_pair = (int(index1), int(index2))
        return self.group_all_distances("__atom1", "__atom2", dim)[:, 0, 0]

    def group_all_distances_reduce(self, group1, group2, dim, cutoff=0.0, out=None):
        """The documented consumer of the matrix (analysis.rs:390-399) fused on the device:
        returns dict(min, argmin [F,2], max, argmax [F,2], count) without materialising the matrix.
        `out`: a dict with those five keys holding host arrays or device tensors (f32 / 2 x u32 / f32 / 2 x u32 / u64
        per frame; torch has no unsigned 64-bit type, int32 pairs / int64 of the same size do) that receive the results;
        with device tensors the call is asynchronous on the ctx's stream."""
        g1, g2 = self._gid(group1), self._gid(group2)
        F = self.n_frames
        r = out if out is not None else {
            "min": np.empty(F, np.float32), "argmin": np.empty((F, 2), np.uint32), "max": np.empty(F, np.float32),
            "argmax": np.empty((F, 2), np.uint32), "count": np.empty(F, np.uint64)}
        self._check(self._lib.groan_gpu_all_distances_reduce(self._h, g1, g2, int(dim), float(cutoff), _ptr(r["min"]),
                                                             _ptr(r["argmin"]), _ptr(r["max"]), _ptr(r["argmax"]),
                                                             _ptr(r["count"])), "group_all_distances_reduce",
                    "%s/%s" % (group1, group2))
        return r

    # ------------------------------------------------------------------ wrap / translate (src/system/modifying.rs)
    def atoms_wrap(self, shifts=False):
        """modifying.rs:201"""
        return self.group_wrap("all", shifts)

    def group_wrap(self, name, shifts=False):
        """modifying.rs:215; shifts=True also returns the int8 image shifts [F, G, 3]"""
        gid = self._gid(name)
        sh = self._shifts_arg(name, shifts)
        self._check(self._lib.groan_gpu_wrap(self._h, gid, _ptr(sh) if sh is not None else None), "group_wrap", name)
        self._bump()
        return sh

    def _shifts_arg(self, name, shifts):
        """shifts: False / None = not wanted, True = a fresh host array, anything else = the caller's int8 buffer
        [F, G, 3] (numpy array or torch tensor, host or device)"""
        if shifts is True:
            return np.empty((self.n_frames, self.group_get_n_atoms(name), 3), np.int8)
        return None if shifts is False or shifts is None else shifts

    def atoms_translate(self, t, shifts=False):
        """modifying.rs:73"""
        return self.group_translate("all", t, shifts)

    def group_translate(self, name, t, shifts=False):
        gid = self._gid(name)
        t3 = (C.c_float * 3)(*[float(v) for v in t])
        sh = self._shifts_arg(name, shifts)
        self._check(self._lib.groan_gpu_translate(self._h, gid, t3, _ptr(sh) if sh is not None else None), "group_translate", name)
        self._bump()
        return sh

    # ------------------------------------------------------------------ cutoff pair search
    def group_pairs_within(self, g1, g2, cutoff, capacity=0, with_distances=False, pairs_out=None, dist_out=None):
        """CellGrid::new(g2, cell >= cutoff) + neighbors_iter around every atom of g1 + distance filter (cellgrid.rs:301-420):
        per frame the number of pairs closer than `cutoff` and, if capacity > 0, up to `capacity` of them as positions inside
        (g1, g2) -- in no particular order -- with their distances.  Returns (count [F], pairs [F, capacity, 2], dist).
        `pairs_out` / `dist_out`: device tensors ([F, capacity, 2] 32-bit integers, [F, capacity] f32) that receive the
        pairs instead of freshly allocated host arrays (the lists stay on the device for whatever consumes them)."""
        F = self.n_frames
        count = np.zeros(F, np.uint64)
        if pairs_out is not None:
            capacity = int(pairs_out.shape[1])
        pairs = pairs_out if pairs_out is not None else (np.zeros((F, capacity, 2), np.uint32) if capacity else None)
        dist = dist_out if dist_out is not None else (np.zeros((F, capacity), np.float32) if (capacity and with_distances) else None)
        self._check(self._lib.groan_gpu_pairs_within(self._h, self._gid(g1), self._gid(g2), C.c_float(cutoff), _ptr(count), _ptr(pairs),
                                                     _ptr(dist), capacity), "group_pairs_within", g1)
        return count, pairs, dist

    # ------------------------------------------------------------------ users of the cell grid
    def guess_bonds(self, vdw, radius_factor=0.55, capacity=None, assign=True):
        """System::guess_bonds (guess.rs:362-395): per frame the bonds i < j with distance(i, j) < (vdw[i] + vdw[j]) *
        radius_factor, found through a cell grid of the longest possible bond.  `vdw`: one van der Waals radius per atom, None /
        NaN / negative = the atom has none (the reference's `no_vdw` warning list: returned as 1-based atom numbers, like
        BondsGuessInfo).  Returns (bonds, no_vdw) with bonds = one sorted [n, 2] array per frame; with `assign`, the bonds of
        frame 0 become the System's topology (assign_bonds + reset_mol_references, guess.rs:409-425)."""
        v = np.array([-1.0 if (x is None or not (x == x) or x < 0) else float(x) for x in vdw], np.float32)
        if v.size != self.n_atoms:
            raise ValueError("one radius per atom")
        F = self.n_frames
        cap = int(capacity) if capacity else 8 * self.n_atoms + 64
        count = np.zeros(F, np.uint64)
        pairs = np.zeros((F, cap, 2), np.uint32)
        self._check(self._lib.groan_gpu_guess_bonds(self._h, _ptr(v), C.c_float(radius_factor), _ptr(count), _ptr(pairs), cap), "guess_bonds")
        if int(count.max()) > cap:
            return self.guess_bonds(vdw, radius_factor, capacity=int(count.max()), assign=assign)
        bonds = []
        for f in range(F):
            p = pairs[f, :int(count[f])].astype(np.int64)
            bonds.append(p[np.lexsort((p[:, 1], p[:, 0]))])
        if assign and F:
            self.add_bonds(bonds[0])
        return bonds, [int(i) + 1 for i in np.nonzero(v < 0)[0]]

    def hbonds_single(self, acceptors, donors, max_distance, min_angle, capacity=None):
        """HBondAnalysis::analyze_single (hbonds.rs:240-320): `acceptors` = a group name, `donors` = [(donor atom, [hydrogen
        atoms])].  Per frame a structured array (donor, hydrogen, acceptor, distance, angle) sorted by (donor, acceptor, hydrogen)."""
        don = np.array([d for d, _ in donors], np.uint32)
        off = np.zeros(len(donors) + 1, np.uint32)
        off[1:] = np.cumsum([len(h) for _, h in donors])
        hyd = np.array([x for _, h in donors for x in h], np.uint32)
        F = self.n_frames
        cap = int(capacity) if capacity else 4 * int(hyd.size) + 64
        count = np.zeros(F, np.uint64)
        dha = np.zeros((F, cap, 3), np.uint32)
        da = np.zeros((F, cap, 2), np.float32)
        self._check(self._lib.groan_gpu_hbonds(self._h, self._gid(acceptors), _ptr(don), _ptr(off), _ptr(hyd if hyd.size else np.zeros(1, np.uint32)),
                                               int(don.size), C.c_float(max_distance), C.c_float(min_angle), _ptr(count), _ptr(dha), _ptr(da), cap),
                    "hbonds", acceptors)
        if int(count.max()) > cap:
            return self.hbonds_single(acceptors, donors, max_distance, min_angle, capacity=int(count.max()))
        out = []
        dt = np.dtype([("donor", np.int64), ("hydrogen", np.int64), ("acceptor", np.int64), ("distance", np.float32), ("angle", np.float32)])
        for f in range(F):
            n = int(count[f])
            r = np.zeros(n, dt)
            r["donor"], r["hydrogen"], r["acceptor"] = dha[f, :n, 0], dha[f, :n, 1], dha[f, :n, 2]
            r["distance"], r["angle"] = da[f, :n, 0], da[f, :n, 1]
            out.append(r[np.lexsort((r["hydrogen"], r["acceptor"], r["donor"]))])
        return out

    # ------------------------------------------------------------------ whole molecules / groups, centering
    def add_bonds(self, pairs):
        """Bonds as (i, j) index pairs, e.g. from CONECT records (System::add_bonds_from_pdb, pdb_io.rs:129-200).  The
        topology stays on the host: molecules = connected components, reference atom = lowest index of a polyatomic
        molecule (System::create_mol_references, modifying.rs:258-283); the device only gets mol_ref[i]."""
        parent = np.arange(self.n_atoms, dtype=np.int64)

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        bonded = np.zeros(self.n_atoms, bool)
        for a, b in np.asarray(pairs, dtype=np.int64).reshape(-1, 2):
            bonded[a] = bonded[b] = True
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        mol_ref = np.full(self.n_atoms, 0xFFFFFFFF, np.uint32)
        for i in np.nonzero(bonded)[0]:
            mol_ref[i] = find(i)
        self._mol_ref = mol_ref
        self._bonds = getattr(self, "_bonds", set()) | {(int(min(a, b)), int(max(a, b))) for a, b in np.asarray(pairs, dtype=np.int64).reshape(-1, 2)}
        self._check(self._lib.groan_gpu_set_molecules(self._h, _ptr(mol_ref)), "add_bonds")

    def make_molecules_whole(self):
        """System::make_molecules_whole (modifying.rs:338-391)"""
        if getattr(self, "_mol_ref", None) is None:  # no bonds: no polyatomic molecule, nothing moves
            self._mol_ref = np.full(self.n_atoms, 0xFFFFFFFF, np.uint32)
            self._check(self._lib.groan_gpu_set_molecules(self._h, _ptr(self._mol_ref)), "make_molecules_whole")
        self._check(self._lib.groan_gpu_make_molecules_whole(self._h), "make_molecules_whole")
        self._bump()

    def make_group_whole(self, name):
        """System::make_group_whole (modifying.rs:437-465)"""
        self._check(self._lib.groan_gpu_make_group_whole(self._h, self._gid(name)), "make_group_whole", name)
        self._bump()

    def atoms_center(self, reference, dimension=Dimension.XYZ):
        """System::atoms_center (utility.rs:109-130)"""
        self._check(self._lib.groan_gpu_atoms_center(self._h, self._gid(reference), 0, int(dimension)), "atoms_center", reference)
        self._bump()

    def atoms_center_mass(self, reference, dimension=Dimension.XYZ):
        """System::atoms_center_mass (utility.rs:168-189)"""
        self._check(self._lib.groan_gpu_atoms_center(self._h, self._gid(reference), 1, int(dimension)), "atoms_center_mass",
                    reference)
        self._bump()

    # ------------------------------------------------------------------ RMSD (src/system/rmsd.rs)
    def _frame0(self):
        if self._host_frames is not None:
            return self._host_frames[0]
        return self.get_frames()[0]

    def _set_reference(self, reference, group):
        gid = self._gid(group, rmsd=True)
        hit = self._ref_cache.get(gid)
        if hit is not None and hit[0]() is reference and hit[1] == reference._version and hit[2] == group:
            return
        # extract_data_from_system(reference) runs first (rmsd.rs:146-147): box, then the group in the REFERENCE
        if reference._boxes is None:
            raise RMSDError("InvalidSimBox(SimBoxError::DoesNotExist)", "reference has no box", _lib.ENOBOX)
        rg = reference._group(group, rmsd=True)
        rmass = getattr(rg, "mass", None)
        if rmass is None and (rg.indices is None or len(rg)):
            first = 0 if rg.indices is None else int(rg.indices[0])
            raise RMSDError("InvalidMass(MassError::NoMass(%d))" % first, "reference atom has no mass", _lib.ENOMASS)
        n_ref = reference.n_atoms if rg.indices is None else len(rg)
        ref_xyz = np.ascontiguousarray(reference._frame0(), dtype=np.float32)
        ref_box = np.ascontiguousarray(reference._boxes[0], dtype=np.float32)
        self._check(self._lib.groan_gpu_rmsd_set_reference(self._h, gid, _ptr(ref_xyz), reference.n_atoms,
                                                           _ptr(rg.indices) if n_ref and rg.indices is not None else None, n_ref,
                                                           _ptr(ref_box), _ptr(rmass) if n_ref else None),
                    "rmsd_set_reference", group, rmsd=True)
        self._ref_cache[gid] = (weakref.ref(reference), reference._version, group)

    def calc_rmsd(self, reference, group, out=None, rot=None):
        """System::calc_rmsd / RMSDTrajRead::calc_rmsd (rmsd.rs:75,315): RMSD of every frame to `reference` (its frame 0)."""
        self._set_reference(reference, group)
        out = self._out(out, (self.n_frames,))
        self._check(self._lib.groan_gpu_rmsd(self._h, self._gid(group, True), _ptr(out), _ptr(rot) if rot is not None else None),
                    "calc_rmsd", group, rmsd=True)
        return out

    def group_center_and_rmsd(self, reference, group, weighted=False, center_out=None, rmsd_out=None, rot=None):
        """group_get_center (or group_get_com when weighted) AND calc_rmsd of `group` from one read of every frame;
        same results as calling the two methods one after the other.  Returns (centre [F,3], rmsd [F])."""
        self._set_reference(reference, group)
        center_out = self._out(center_out, (self.n_frames, 3))
        rmsd_out = self._out(rmsd_out, (self.n_frames,))
        self._check(self._lib.groan_gpu_center_rmsd(self._h, self._gid(group, True), 1 if weighted else 0, _ptr(center_out),
                                                    _ptr(rmsd_out), _ptr(rot) if rot is not None else None),
                    "group_center_and_rmsd", group, rmsd=True)
        return center_out, rmsd_out

    def calc_rmsd_and_fit(self, reference, group, out=None):
        """System::calc_rmsd_and_fit (rmsd.rs:129; fit_structure :508-528): also fits ALL atoms of every frame in place."""
        self._set_reference(reference, group)
        out = self._out(out, (self.n_frames,))
        self._check(self._lib.groan_gpu_rmsd_fit(self._h, self._gid(group, True), _ptr(out)), "calc_rmsd_and_fit", group, rmsd=True)
        self._bump()
        return out
