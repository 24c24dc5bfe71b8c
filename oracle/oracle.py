"""ctypes front end of the CPU oracle (oracle/groan_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (groan_rs_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, ENOBOX, ENOTORTHO, EEMPTY, ENOPOS, ENOMASS, EGROUPSIZE, EZEROBOX = range(8)
DIM = {"None": 0, "X": 1, "Y": 2, "Z": 3, "XY": 4, "XZ": 5, "YZ": 6, "XYZ": 7}

_f = C.POINTER(C.c_float)
_d = C.POINTER(C.c_double)
_u = C.POINTER(C.c_uint32)
_i8 = C.POINTER(C.c_int8)
_sz = C.c_size_t


def build(force=False):
    so = os.path.join(HERE, "libgroan_oracle.so")
    src = os.path.join(HERE, "groan_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libgroan_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_wrap1.restype = C.c_float
        L.orc_wrap1.argtypes = [C.c_float, C.c_float]
        L.orc_minimg1.restype = C.c_float
        L.orc_minimg1.argtypes = [C.c_float, C.c_float]
        L.orc_floor_mod.restype = C.c_float
        L.orc_floor_mod.argtypes = [C.c_float, C.c_float]
        L.orc_distance.restype = C.c_float
        L.orc_distance.argtypes = [_f, _f, C.c_int, _f]
        L.orc_tric_distance.restype = C.c_float
        L.orc_tric_distance.argtypes = [_f, _f, C.c_int, _f]
        L.orc_tric_distance_brute64.restype = C.c_double
        L.orc_tric_distance_brute64.argtypes = [_f, _f, C.c_int, _f, C.c_int]
        L.orc_hash.restype = C.c_uint64
        L.orc_hash.argtypes = [C.c_uint64] * 4
        L.orc_baseline_traj.restype = C.c_double
        L.orc_baseline_pairs.restype = C.c_double
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(_f)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _idx(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _L(box):
    """accepts L[3] or a 3x3 / 9 matrix; returns (status, L[3])"""
    b = _f32(box).reshape(-1)
    if b.size == 3:
        return OK, b.copy()
    out = np.zeros(3, np.float32)
    st = lib().orc_box_lengths(_fp(b), _fp(out))
    return st, out


class OracleError(Exception):
    def __init__(self, code):
        super().__init__("oracle status %d" % code)
        self.code = code


def _chk(st):
    if st != OK:
        raise OracleError(st)


def wrap1(x, L):
    return lib().orc_wrap1(np.float32(x), np.float32(L))


def minimg1(d, L):
    return lib().orc_minimg1(np.float32(d), np.float32(L))


def floor_mod(x, y):
    return lib().orc_floor_mod(np.float32(x), np.float32(y))


def vector_to(c, p, box):
    st, L = _L(box)
    _chk(st)
    out = np.zeros(3, np.float32)
    lib().orc_vector_to(_fp(_f32(c)), _fp(_f32(p)), _fp(L), _fp(out))
    return out


def distance(a, b, dim, box):
    st, L = _L(box)
    _chk(st)
    return np.float32(lib().orc_distance(_fp(_f32(a)), _fp(_f32(b)), DIM[dim], _fp(L)))


def _grp(xyz, idx):
    x = _f32(xyz).reshape(-1, 3)
    i = _idx(idx)
    return x, i


def estimate_center(xyz, idx, box, mass=None):
    st, L = _L(box)
    _chk(st)
    x, i = _grp(xyz, idx)
    out = np.zeros(3, np.float32)
    m = _f32(mass) if mass is not None else None
    _chk(lib().orc_estimate_center(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(m) if m is not None else None,
                                   _fp(L), _fp(out)))
    return out


def get_center(xyz, idx, box):
    st, L = _L(box)
    _chk(st)
    x, i = _grp(xyz, idx)
    if i.size == 0:
        raise OracleError(EEMPTY)
    out = np.zeros(3, np.float32)
    _chk(lib().orc_get_center(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(L), _fp(out)))
    return out


def get_com(xyz, idx, mass, box):
    st, L = _L(box)
    _chk(st)
    x, i = _grp(xyz, idx)
    if i.size == 0:
        raise OracleError(EEMPTY)
    out = np.zeros(3, np.float32)
    m = _f32(mass)
    _chk(lib().orc_get_com(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(m), _fp(L), _fp(out)))
    return out


def get_center_naive(xyz, idx):
    x, i = _grp(xyz, idx)
    out = np.zeros(3, np.float32)
    lib().orc_get_center_naive(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(out))
    return out


def group_distance(xyz, idx1, idx2, dim, box):
    st, L = _L(box)
    _chk(st)
    x, i1 = _grp(xyz, idx1)
    i2 = _idx(idx2)
    out = C.c_float(0)
    _chk(lib().orc_group_distance(_fp(x), _sz(3), i1.ctypes.data_as(_u), _sz(i1.size), i2.ctypes.data_as(_u), _sz(i2.size),
                                  DIM[dim], _fp(L), C.byref(out)))
    return np.float32(out.value)


def all_distances(xyz, idx1, idx2, dim, box):
    st, L = _L(box)
    _chk(st)
    x, i1 = _grp(xyz, idx1)
    i2 = _idx(idx2)
    out = np.zeros((i1.size, i2.size), np.float32)
    _chk(lib().orc_all_distances(_fp(x), _sz(3), i1.ctypes.data_as(_u), _sz(i1.size), i2.ctypes.data_as(_u), _sz(i2.size),
                                 DIM[dim], _fp(L), _fp(out)))
    return out


def all_distances_minmax(xyz, idx1, idx2, dim, box, cutoff=0.0):
    st, L = _L(box)
    _chk(st)
    x, i1 = _grp(xyz, idx1)
    i2 = _idx(idx2)
    dmin, dmax = C.c_float(0), C.c_float(0)
    imin, imax = np.zeros(2, np.uint32), np.zeros(2, np.uint32)
    cnt = C.c_uint64(0)
    _chk(lib().orc_all_distances_minmax(_fp(x), _sz(3), i1.ctypes.data_as(_u), _sz(i1.size), i2.ctypes.data_as(_u),
                                        _sz(i2.size), DIM[dim], _fp(L), C.byref(dmin), imin.ctypes.data_as(_u),
                                        C.byref(dmax), imax.ctypes.data_as(_u), C.c_float(cutoff), C.byref(cnt)))
    return np.float32(dmin.value), imin, np.float32(dmax.value), imax, cnt.value


def wrap(xyz, idx, box):
    """returns (wrapped copy, shifts int8 [g,3])"""
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3).copy()
    i = _idx(idx)
    sh = np.zeros((i.size, 3), np.int8)
    _chk(lib().orc_wrap(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(L), sh.ctypes.data_as(_i8)))
    return x, sh


def translate(xyz, idx, t, box):
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3).copy()
    i = _idx(idx)
    sh = np.zeros((i.size, 3), np.int8)
    _chk(lib().orc_translate(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(_f32(t)), _fp(L), sh.ctypes.data_as(_i8)))
    return x, sh


def pairs_within(xyz, idx1, idx2, cutoff, box, capacity=0):
    """brute-force restatement of CellGrid::neighbors_iter + distance filter: (count, pairs [n,2], dist [n]) row-major"""
    st, L = _L(box)
    _chk(st)
    x, i1 = _grp(xyz, idx1)
    i2 = _idx(idx2)
    cnt = C.c_uint64(0)
    pairs = np.zeros((max(capacity, 1), 2), np.uint32)
    dist = np.zeros(max(capacity, 1), np.float32)
    _chk(lib().orc_pairs_within(_fp(x), _sz(3), i1.ctypes.data_as(_u), _sz(i1.size), i2.ctypes.data_as(_u), _sz(i2.size),
                                C.c_float(cutoff), _fp(L), C.byref(cnt), pairs.ctypes.data_as(_u) if capacity else None,
                                _fp(dist) if capacity else None, _sz(capacity)))
    n = min(int(cnt.value), capacity)
    return int(cnt.value), pairs[:n], dist[:n]


NO_MOL = 0xFFFFFFFF


def make_group_whole(xyz, idx, box):
    """System::make_group_whole (modifying.rs:437-465); returns the modified copy"""
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3).copy()
    i = _idx(idx)
    _chk(lib().orc_make_group_whole(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(L)))
    return x


def make_molecules_whole(xyz, mol_ref, box):
    """System::make_molecules_whole (modifying.rs:338-391); mol_ref[i] = reference atom of i's molecule or NO_MOL"""
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3).copy()
    r = np.ascontiguousarray(mol_ref, dtype=np.uint32)
    _chk(lib().orc_make_molecules_whole(_fp(x), _sz(3), _sz(x.shape[0]), r.ctypes.data_as(_u), _fp(L)))
    return x


def atoms_center(xyz, idx, dim, box, mass=None):
    """System::atoms_center / atoms_center_mass (utility.rs:109-189); dim as in distance(); returns the modified copy"""
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3).copy()
    i = _idx(idx)
    m = _f32(mass) if mass is not None else None
    _chk(lib().orc_atoms_center(_fp(x), _sz(3), _sz(x.shape[0]), i.ctypes.data_as(_u), _sz(i.size),
                                _fp(m) if m is not None else None, C.c_int(DIM[dim]), _fp(L)))
    return x


def kabsch(p, q, w, cp, cq, sum_w):
    p, q, w = _f32(p).reshape(-1, 3), _f32(q).reshape(-1, 3), _f32(w)
    r, t = np.zeros(9, np.float32), np.zeros(3, np.float32)
    rmsd = C.c_float(0)
    lib().orc_kabsch(_fp(p), _fp(q), _fp(w), _sz(p.shape[0]), _fp(_f32(cp)), _fp(_f32(cq)), C.c_float(sum_w), _fp(r), _fp(t),
                     C.byref(rmsd))
    return r.reshape(3, 3), t, np.float32(rmsd.value)


def calc_rmsd(ref_xyz, ref_idx, ref_box, mass, tgt_xyz, tgt_idx, tgt_box):
    """returns (rmsd f32, r[3,3] row-major)"""
    st, Lr = _L(ref_box)
    _chk(st)
    st, Lt = _L(tgt_box)
    _chk(st)
    rx, ri = _grp(ref_xyz, ref_idx)
    tx, ti = _grp(tgt_xyz, tgt_idx)
    if mass is None:
        raise OracleError(ENOMASS)
    m = _f32(mass)
    r = np.zeros(9, np.float32)
    rmsd = C.c_float(0)
    _chk(lib().orc_calc_rmsd(_fp(rx), _sz(3), ri.ctypes.data_as(_u), _sz(ri.size), _fp(Lr), _fp(m), _fp(tx), _sz(3),
                             ti.ctypes.data_as(_u), _sz(ti.size), _fp(Lt), _fp(r), C.byref(rmsd)))
    return np.float32(rmsd.value), r.reshape(3, 3)


def fit(xyz, r, com_tgt, com_ref, box):
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3).copy()
    lib().orc_fit(_fp(x), _sz(3), _sz(x.shape[0]), _fp(_f32(r).reshape(-1)), _fp(_f32(com_tgt)), _fp(_f32(com_ref)), _fp(L))
    return x


def calc_rmsd_and_fit(ref_xyz, ref_idx, ref_box, mass, tgt_xyz, tgt_idx, tgt_box):
    """rmsd.rs:129-138: returns (rmsd, fitted copy of all target atoms)"""
    rmsd, r = calc_rmsd(ref_xyz, ref_idx, ref_box, mass, tgt_xyz, tgt_idx, tgt_box)
    com_ref = get_com(ref_xyz, ref_idx, mass, ref_box)
    com_tgt = get_com(tgt_xyz, tgt_idx, mass, tgt_box)
    return rmsd, fit(tgt_xyz, r, com_tgt, com_ref, tgt_box)


# ---------------------------------------------------------------- exact64


def estimate_center_x64(xyz, idx, box, mass=None):
    st, L = _L(box)
    _chk(st)
    x, i = _grp(xyz, idx)
    out = np.zeros(3, np.float64)
    m = _f32(mass) if mass is not None else None
    _chk(lib().orc_estimate_center_x64(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(m) if m is not None else None,
                                       _fp(L), out.ctypes.data_as(_d)))
    return out


def get_center_x64(xyz, idx, box, mass=None):
    st, L = _L(box)
    _chk(st)
    x, i = _grp(xyz, idx)
    out = np.zeros(3, np.float64)
    m = _f32(mass) if mass is not None else None
    _chk(lib().orc_get_center_x64(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(m) if m is not None else None,
                                  _fp(L), out.ctypes.data_as(_d)))
    return out


def calc_rmsd_x64(ref_xyz, ref_idx, ref_box, mass, tgt_xyz, tgt_idx, tgt_box):
    st, Lr = _L(ref_box)
    _chk(st)
    st, Lt = _L(tgt_box)
    _chk(st)
    rx, ri = _grp(ref_xyz, ref_idx)
    tx, ti = _grp(tgt_xyz, tgt_idx)
    m = _f32(mass)
    r = np.zeros(9, np.float64)
    rmsd = C.c_double(0)
    _chk(lib().orc_calc_rmsd_x64(_fp(rx), _sz(3), ri.ctypes.data_as(_u), _sz(ri.size), _fp(Lr), _fp(m), _fp(tx), _sz(3),
                                 ti.ctypes.data_as(_u), _sz(ti.size), _fp(Lt), r.ctypes.data_as(_d), C.byref(rmsd)))
    return rmsd.value, r.reshape(3, 3)


# ---------------------------------------------------------------- triclinic extension (UNPINNED by the reference)


def tric_wrap(xyz, idx, box9):
    x = _f32(xyz).reshape(-1, 3).copy()
    i = _idx(idx)
    b = _f32(box9).reshape(-1)
    sh = np.zeros((i.size, 3), np.int8)
    _chk(lib().orc_tric_wrap(_fp(x), _sz(3), i.ctypes.data_as(_u), _sz(i.size), _fp(b), sh.ctypes.data_as(_i8)))
    return x, sh


def tric_distance(a, b, dim, box9):
    return np.float32(lib().orc_tric_distance(_fp(_f32(a)), _fp(_f32(b)), DIM[dim], _fp(_f32(box9).reshape(-1))))


def tric_distance_brute64(a, b, dim, box9, nimg=2):
    return lib().orc_tric_distance_brute64(_fp(_f32(a)), _fp(_f32(b)), DIM[dim], _fp(_f32(box9).reshape(-1)), nimg)


def tric_all_distances(xyz, idx1, idx2, dim, box9):
    x, i1 = _grp(xyz, idx1)
    i2 = _idx(idx2)
    out = np.zeros((i1.size, i2.size), np.float32)
    _chk(lib().orc_tric_all_distances(_fp(x), _sz(3), i1.ctypes.data_as(_u), _sz(i1.size), i2.ctypes.data_as(_u), _sz(i2.size),
                                      DIM[dim], _fp(_f32(box9).reshape(-1)), _fp(out)))
    return out



# Triclinic centre / RMSD: the DEFINITION of the extension ("Bai-Breen on fractional coordinates, mapped back"),
# restated in numpy float64.  A box v1 = (a, 0, 0), v2 = (bx, by, 0), v3 = (cx, cy, cz) becomes the orthogonal periodic box
# (a, by, cz) under the shear u_z = z, u_y = y - z cy/cz, u_x = x - u_y bx/by - z cx/cz; every orthogonal algorithm of this
# file is applied to u and the result mapped back (the map is linear).  Pins: tests/test_gpu_parity.py
# test_triclinic_centre_and_rmsd_* (brute-force nearest-image unwrapping of compact groups; identity on orthogonal boxes).
def _tric_params(box9):
    b = np.asarray(box9, np.float64).reshape(3, 3)
    L = np.array([b[0, 0], b[1, 1], b[2, 2]])
    return b[1, 0] / b[1, 1], b[2, 0] / b[2, 2], b[2, 1] / b[2, 2], L


def tric_to_u(x, box9):
    a21, a31, a32, _ = _tric_params(box9)
    x = np.asarray(x, np.float64).reshape(-1, 3)
    uy = x[:, 1] - x[:, 2] * a32
    ux = x[:, 0] - uy * a21 - x[:, 2] * a31
    return np.stack([ux, uy, x[:, 2]], axis=1)


def tric_to_x(u, box9):
    a21, a31, a32, _ = _tric_params(box9)
    u = np.asarray(u, np.float64).reshape(-1, 3)
    return np.stack([u[:, 0] + u[:, 1] * a21 + u[:, 2] * a31, u[:, 1] + u[:, 2] * a32, u[:, 2]], axis=1)


def _center_u64(u, L, mass=None):
    th = 2.0 * np.pi * np.mod(u, L) / L
    c0 = (np.arctan2(-np.sin(th).sum(axis=0), -np.cos(th).sum(axis=0)) + np.pi) * L / (2.0 * np.pi)  # geometric estimate
    v = np.mod(u - c0 + L / 2.0, L) - L / 2.0
    w = np.ones(len(u)) if mass is None else np.asarray(mass, np.float64)
    return ((c0 + v) * w[:, None]).sum(axis=0) / w.sum(), c0


def tric_estimate_center_x64(xyz, idx, box9):
    u = tric_to_u(np.asarray(xyz, np.float64).reshape(-1, 3)[np.asarray(idx, np.int64)], box9)
    return tric_to_x(_center_u64(u, _tric_params(box9)[3])[1], box9)[0]


def tric_get_center_x64(xyz, idx, box9, mass=None):
    u = tric_to_u(np.asarray(xyz, np.float64).reshape(-1, 3)[np.asarray(idx, np.int64)], box9)
    return tric_to_x(_center_u64(u, _tric_params(box9)[3], mass)[0], box9)[0]


def _tric_extract(xyz, idx, box9, mass):
    L = _tric_params(box9)[3]
    u = tric_to_u(np.asarray(xyz, np.float64).reshape(-1, 3)[np.asarray(idx, np.int64)], box9)
    com = _center_u64(u, L, mass)[0]
    return tric_to_x(np.mod(u + (L / 2.0 - com), L) - L / 2.0, box9)  # wrapped around the COM, centred, Cartesian


def tric_calc_rmsd_x64(ref_xyz, ref_idx, ref_box9, mass, tgt_xyz, tgt_idx, tgt_box9):
    pc = _tric_extract(ref_xyz, ref_idx, ref_box9, mass)
    qc = _tric_extract(tgt_xyz, tgt_idx, tgt_box9, mass)
    w = np.asarray(mass, np.float64)
    U, _, Vt = np.linalg.svd(pc.T @ qc)
    D = np.diag([1.0, 1.0, -1.0 if np.linalg.det(U @ Vt) < 0 else 1.0])
    r = U @ D @ Vt
    diff = pc @ r - qc  # r^T pc_i as rows
    return float(np.sqrt((w * (diff * diff).sum(axis=1)).sum() / w.sum())), r



# ---------------------------------------------------------------- users of the cell grid (guess.rs:362-470, hbonds.rs:240-335)
# numpy float32 restatements with the reference's operation order; brute force over all pairs instead of the cell grid (the
# grid only prunes: cells are at least as wide as the cutoff, so neighbors_iter offers every atom within it).
def _minimg_vec(d, L):
    """min_image per axis with the reference's loops (vector3d.rs:575-592), vectorised, float32"""
    d = d.astype(np.float32).copy()
    L = np.asarray(L, np.float32)
    h = (L / np.float32(2.0)).astype(np.float32)
    for _ in range(64):
        hi = d > h
        if not hi.any():
            break
        d = np.where(hi, (d - L).astype(np.float32), d)
    for _ in range(64):
        lo = d < -h
        if not lo.any():
            break
        d = np.where(lo, (d + L).astype(np.float32), d)
    return d


def _dist32(a, b, L):
    """Vector3D::distance, Dimension::XYZ: sqrt((dx*dx + dy*dy) + dz*dz) in float32"""
    d = _minimg_vec((np.asarray(a, np.float32) - np.asarray(b, np.float32)).astype(np.float32), L)
    q = ((d[..., 0] * d[..., 0]).astype(np.float32) + (d[..., 1] * d[..., 1]).astype(np.float32)).astype(np.float32)
    q = (q + (d[..., 2] * d[..., 2]).astype(np.float32)).astype(np.float32)
    return np.sqrt(q).astype(np.float32)


def guess_bonds(xyz, vdw, box, radius_factor=0.55):
    """identify_bonds (guess.rs:427-470): sorted array of (i, j), i < j; vdw < 0 / NaN = no radius"""
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3)
    v = np.asarray(vdw, np.float32)
    has = v >= 0
    out = []
    rf = np.float32(radius_factor)
    for i in np.nonzero(has)[0]:
        js = np.nonzero(has & (np.arange(len(v)) > i))[0]
        if js.size == 0:
            continue
        d = _dist32(x[i][None, :], x[js], L)
        lim = ((v[i] + v[js]).astype(np.float32) * rf).astype(np.float32)
        for j in js[d < lim]:
            out.append((int(i), int(j)))
    return np.array(sorted(out), np.int64).reshape(-1, 2)


def _vector_to32(c, p, L):
    L = np.asarray(L, np.float32)
    h = (L / np.float32(2.0)).astype(np.float32)
    t = ((np.asarray(p, np.float32) - np.asarray(c, np.float32)).astype(np.float32) + h).astype(np.float32)
    m = np.fmod((np.fmod(t, L).astype(np.float32) + L).astype(np.float32), L).astype(np.float32)
    return (m - h).astype(np.float32)


def hbond_angle(donor, hydrogen, acceptor, box):
    """HBondAnalysis::calc_angle (hbonds.rs:322-350), degrees, float32"""
    st, L = _L(box)
    _chk(st)
    hd, ha = _vector_to32(hydrogen, donor, L), _vector_to32(hydrogen, acceptor, L)
    dot = np.float32(np.float32(np.float32(hd[0] * ha[0]) + np.float32(hd[1] * ha[1])) + np.float32(hd[2] * ha[2]))
    n1 = np.sqrt(np.float32(np.float32(np.float32(hd[0] * hd[0]) + np.float32(hd[1] * hd[1])) + np.float32(hd[2] * hd[2])))
    n2 = np.sqrt(np.float32(np.float32(np.float32(ha[0] * ha[0]) + np.float32(ha[1] * ha[1])) + np.float32(ha[2] * ha[2])))
    with np.errstate(invalid="ignore", divide="ignore"):
        ang = np.float32(np.arccos(np.float32(dot / np.float32(n1 * n2))) * np.float32(57.29577951308232))
    if np.isnan(ang):
        return np.float32(180.0) if _dist32(hydrogen, acceptor, L) < _dist32(donor, acceptor, L) else np.float32(0.0)
    return ang


def hbonds_single(xyz, acceptors, donors, box, max_distance, min_angle):
    """HBondAnalysis::analyze_single (hbonds.rs:240-320): list of (donor, hydrogen, acceptor, distance, angle), sorted by
    (donor, acceptor, hydrogen); donors = [(donor, [hydrogens])]"""
    st, L = _L(box)
    _chk(st)
    x = _f32(xyz).reshape(-1, 3)
    acc = np.asarray(acceptors, np.int64)
    out = []
    for d, hs in donors:
        dist = _dist32(x[acc], x[d][None, :], L)
        for a, dd in zip(acc[(dist <= np.float32(max_distance)) & (acc != d)], dist[(dist <= np.float32(max_distance)) & (acc != d)]):
            for h in hs:
                ang = hbond_angle(x[d], x[h], x[a], L)
                if ang < np.float32(min_angle):
                    continue
                out.append((int(d), int(h), int(a), float(dd), float(ang)))
    return sorted(out, key=lambda r: (r[0], r[2], r[1]))


# ---------------------------------------------------------------- synthetic workloads


def synth_uniform(n, seed, frame, lo, span):
    x = np.zeros((n, 3), np.float32)
    lib().orc_synth_uniform(_fp(x), _sz(n), C.c_uint64(seed), C.c_uint64(frame), _fp(_f32(lo)), _fp(_f32(span)))
    return x


def synth_blob_ref(n, seed, scale, centre):
    x = np.zeros((n, 3), np.float32)
    lib().orc_synth_blob_ref(_fp(x), _sz(n), C.c_uint64(seed), C.c_float(scale), _fp(_f32(centre)))
    return x


def synth_blob_frame(n, seed, frame, scale, nscale, rot, centre, L, wrap=True):
    x = np.zeros((n, 3), np.float32)
    lib().orc_synth_blob_frame(_fp(x), _sz(n), C.c_uint64(seed), C.c_uint64(frame), C.c_float(scale), C.c_float(nscale),
                               _fp(_f32(rot).reshape(-1)), _fp(_f32(centre)), _fp(_f32(L)), C.c_int(1 if wrap else 0))
    return x


# ---------------------------------------------------------------- restated CPU trajectory path (baseline)


def baseline_traj(frames, boxes_L, idx, mass_all, ref_xyz, ref_L, ops, n_threads, total_frames=None):
    """ops bitmask: 1 group_get_center, 2 calc_rmsd, 4 fit, 8 atoms_wrap.  Returns (seconds, centers, rmsd).
    total_frames: length of the trajectory to process; the stored frames are visited cyclically (frame f = frames[f % F])."""
    fr = _f32(frames)
    F, n = fr.shape[0], fr.shape[1]
    if total_frames is not None:
        bx = _f32(boxes_L).reshape(F, 3)
        i = _idx(idx)
        cen, rm = np.zeros((F, 3), np.float32), np.zeros(F, np.float32)
        m = _f32(mass_all) if mass_all is not None else None
        rx = _f32(ref_xyz) if ref_xyz is not None else None
        rl = _f32(ref_L) if ref_L is not None else None
        fn = lib().orc_baseline_traj_cyclic
        fn.restype = C.c_double
        sec = fn(_fp(fr), _fp(bx), _sz(F), _sz(int(total_frames)), _sz(n), i.ctypes.data_as(_u), _sz(i.size),
                 _fp(m) if m is not None else None, _fp(rx) if rx is not None else None, _fp(rl) if rl is not None else None,
                 C.c_int(ops), C.c_int(n_threads), _fp(cen), _fp(rm))
        return sec, cen, rm
    bx = _f32(boxes_L).reshape(F, 3)
    i = _idx(idx)
    cen = np.zeros((F, 3), np.float32)
    rm = np.zeros(F, np.float32)
    m = _f32(mass_all) if mass_all is not None else None
    rx = _f32(ref_xyz) if ref_xyz is not None else None
    rl = _f32(ref_L) if ref_L is not None else None
    sec = lib().orc_baseline_traj(_fp(fr), _fp(bx), _sz(F), _sz(n), i.ctypes.data_as(_u), _sz(i.size),
                                  _fp(m) if m is not None else None, _fp(rx) if rx is not None else None,
                                  _fp(rl) if rl is not None else None, C.c_int(ops), C.c_int(n_threads), _fp(cen), _fp(rm))
    return sec, cen, rm


def baseline_pairs(frames, boxes_L, idx1, idx2, dim, n_threads):
    fr = _f32(frames)
    F, n = fr.shape[0], fr.shape[1]
    bx = _f32(boxes_L).reshape(F, 3)
    i1, i2 = _idx(idx1), _idx(idx2)
    dmin, dmax = np.zeros(F, np.float32), np.zeros(F, np.float32)
    sec = lib().orc_baseline_pairs(_fp(fr), _fp(bx), _sz(F), _sz(n), i1.ctypes.data_as(_u), _sz(i1.size),
                                   i2.ctypes.data_as(_u), _sz(i2.size), C.c_int(DIM[dim]), C.c_int(n_threads), _fp(dmin),
                                   _fp(dmax))
    return sec, dmin, dmax
