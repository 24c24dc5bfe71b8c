#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own fixtures (TEST INFRASTRUCTURE).

Runs only in the build container (needs /root/reference and oracle/_ref/libxdrfile.so,
built by `make -C oracle ref` from the reference's vendored external/xdrfile C sources where
they lie).  The GPU box has no /root/reference, so the decoded inputs and the reference's
golden outputs travel as small .npz files under tests/golden/.

Parsing rules follow the reference so the decoded numbers line up:
  GRO  fixed columns           src/io/gro_io/structure.rs:120-231, box line src/io/gro_io/mod.rs
  NDX  1-based, sorted+dedup   src/io/ndx_io.rs:103-228, src/structures/container.rs:51-115
  XTC  read_xtc, box[i][j]     external/xdrfile/xdrfile_xtc.h:45-52, src/io/xdrfile.rs:170-199
  @protein / @membrane macros  src/select/mod.rs:591-620
"""
import ctypes
import os
import struct
import sys

import numpy as np

REF = os.environ.get("GROAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TF = os.path.join(REF, "test_files")


def read_gro(path):
    with open(path) as fh:
        lines = fh.read().split("\n")
    n = int(lines[1].split()[0])
    xyz = np.zeros((n, 3), np.float32)
    resname, atomname = [], []
    for i in range(n):
        ln = lines[2 + i]
        resname.append(ln[5:10].strip())
        atomname.append(ln[10:15].strip())
        xyz[i] = [np.float32(ln[20:28]), np.float32(ln[28:36]), np.float32(ln[36:44])]
    b = [np.float32(t) for t in lines[2 + n].split()]
    # gro order: v1x v2y v3z v1y v1z v2x v2z v3x v3y  -> row-major matrix rows v1,v2,v3
    b = b + [np.float32(0)] * (9 - len(b))
    box = np.array([[b[0], b[3], b[4]], [b[5], b[1], b[6]], [b[7], b[8], b[2]]], np.float32)
    return xyz, box, resname, atomname


def read_ndx(path):
    groups, cur = {}, None
    with open(path) as fh:
        for ln in fh:
            ln = ln.strip()
            if ln.startswith("["):
                cur = ln.strip("[] ")
                groups[cur] = []
            elif ln and cur is not None:
                groups[cur] += [int(t) - 1 for t in ln.split()]
    return {k: np.array(sorted(set(v)), np.uint32) for k, v in groups.items()}


class Xdr:
    def __init__(self):
        self.lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libxdrfile.so"))
        self.lib.xdrfile_open.restype = ctypes.c_void_p
        self.lib.xdrfile_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        self.lib.xdrfile_close.argtypes = [ctypes.c_void_p]
        self.lib.read_xtc_natoms.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        self.lib.read_xtc.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.POINTER(ctypes.c_float)]

    def read_xtc(self, path):
        n = ctypes.c_int(0)
        assert self.lib.read_xtc_natoms(path.encode(), ctypes.byref(n)) == 0
        xd = self.lib.xdrfile_open(path.encode(), b"r")
        frames, boxes, times, precision = [], [], [], 0.0
        while True:
            x = np.zeros((n.value, 3), np.float32)
            box = np.zeros((3, 3), np.float32)
            step, time, prec = ctypes.c_int(0), ctypes.c_float(0), ctypes.c_float(0)
            rc = self.lib.read_xtc(xd, n.value, ctypes.byref(step), ctypes.byref(time), box.ctypes.data, x.ctypes.data,
                                   ctypes.byref(prec))
            if rc != 0:
                assert rc == 11, rc  # exdrENDOFFILE  (xdrfile_xtc.rs:63-83)
                break
            frames.append(x)
            boxes.append(box)
            times.append(time.value)
            precision = prec.value
        self.lib.xdrfile_close(xd)
        return np.stack(frames), np.stack(boxes), np.array(times, np.float32), precision


def quantise(frames, prec):
    """xtc coordinates are int * (1/precision) in f32 (external/xdrfile/xdrfile.c decode loop);
    store the ints and prove the floats are reproduced bit for bit."""
    inv = np.float32(1.0) / np.float32(prec)
    q = np.rint(frames.astype(np.float64) * prec).astype(np.int32)
    back = q.astype(np.float32) * inv
    assert np.array_equal(back.view(np.uint32), frames.view(np.uint32)), "quantised xtc round trip not bit exact"
    assert np.abs(q).max() < 32767
    return q.astype(np.int16)


def read_pdb(path):
    """ATOM/HETATM coordinates in nm (f32 parse, then / 10.0 in f32: pdb_io.rs:348-400), CRYST1 lengths (pdb_io.rs:412-430),
    CONECT bonds as index pairs (atom numbers -> indices, pdb_io.rs:129-200,463-520)"""
    xyz, numbers, box, conect = [], [], None, []
    with open(path) as fh:
        for ln in fh:
            ln = ln.rstrip("\n")
            if ln[:4] == "ATOM" or ln[:6] == "HETATM":
                numbers.append(int(ln[6:11]))
                xyz.append([np.float32(ln[30 + 8 * k:38 + 8 * k]) / np.float32(10.0) for k in range(3)])
            elif ln[:6] == "CRYST1":
                box = [np.float32(ln[6 + 9 * k:15 + 9 * k]) / np.float32(10.0) for k in range(3)]
            elif ln[:6] == "CONECT":
                conect.append(ln)
    num2idx = {a: i for i, a in enumerate(numbers)}
    bonds = set()
    for ln in conect:
        a = num2idx[int(ln[6:11])]
        it = 11
        while it + 4 < len(ln):
            t = ln[it:it + 5].strip()
            if t:
                b = num2idx[int(t)]
                if a != b:
                    bonds.add((min(a, b), max(a, b)))
            it += 5
    return np.array(xyz, np.float32), np.array(box, np.float32), np.array(sorted(bonds), np.uint32)


PROTEIN_RES = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLU", "GLN", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE",
               "PRO", "SER", "THR", "TRP", "TYR", "VAL"}


def main():
    os.makedirs(OUT, exist_ok=True)
    xdr = Xdr()

    # ---- example.gro + index.ndx (analysis.rs KATs) and short_trajectory (rmsd.rs KATs)
    xyz, box, _, _ = read_gro(os.path.join(TF, "example.gro"))
    ndx = read_ndx(os.path.join(TF, "index.ndx"))
    np.savez_compressed(os.path.join(OUT, "example.npz"), xyz=xyz, box=box, Protein=ndx["Protein"], Membrane=ndx["Membrane"])

    fr, bx, tm, prec = xdr.read_xtc(os.path.join(TF, "short_trajectory.xtc"))
    fit, fbx, _, fprec = xdr.read_xtc(os.path.join(TF, "short_trajectory_fit.xtc"))
    assert fr.shape == fit.shape == (11, 16844, 3)
    # Martini masses of the 61 protein beads from example.tpr: big-endian f32 records of stride
    # 32 B (mass, charge, massB, chargeB) starting at byte 16834
    raw = open(os.path.join(TF, "example.tpr"), "rb").read()
    masses = np.array([struct.unpack(">f", raw[16834 + 32 * i: 16838 + 32 * i])[0] for i in range(61)], np.float32)
    assert set(masses.tolist()) <= {36.0, 54.0, 72.0}, masses
    np.savez_compressed(os.path.join(OUT, "short_trajectory.npz"), q=quantise(fr, prec), box=bx, time=tm,
                        prec=np.float32(prec), fit_q=quantise(fit, fprec), fit_box=fbx, protein_mass=masses)

    # ---- cfg1: protein.gro + short_trajectory_protein.xtc
    pxyz, pbox, _, _ = read_gro(os.path.join(TF, "protein.gro"))
    pfr, pbx, ptm, pprec = xdr.read_xtc(os.path.join(TF, "short_trajectory_protein.xtc"))
    assert pfr.shape == (11, 61, 3)
    np.savez_compressed(os.path.join(OUT, "protein.npz"), xyz=pxyz, box=pbox, frames=pfr, boxes=pbx, time=ptm, mass=masses)

    # ---- cfg2: aa_membrane_peptide: sub-system = peptide (@protein) + phosphates (@membrane and name P)
    axyz, abox, resn, atn = read_gro(os.path.join(TF, "aa_membrane_peptide.gro"))
    afr, abx, atm, _ = xdr.read_xtc(os.path.join(TF, "aa_membrane_peptide.xtc"))
    pep = np.array([i for i, r in enumerate(resn) if r in PROTEIN_RES], np.uint32)
    pat = np.array([i for i, (r, a) in enumerate(zip(resn, atn)) if r == "POPC" and a == "P"], np.uint32)
    assert len(pep) == 363 and len(pat) == 128 and afr.shape == (21, 32817, 3), (len(pep), len(pat), afr.shape)
    keep = np.concatenate([pep, pat])
    assert np.all(np.diff(keep.astype(np.int64)) > 0)
    np.savez_compressed(os.path.join(OUT, "aa_membrane_peptide.npz"), gro_xyz=axyz[keep], gro_box=abox, frames=afr[:, keep],
                        boxes=abx, time=atm, orig_index=keep, Peptide=np.arange(363, dtype=np.uint32),
                        Phosphates=np.arange(363, 491, dtype=np.uint32))

    # ---- the water of the same system for the hbonds.rs goldens (hbonds.rs:505-587): indices only -- the
    # trajectory itself is tests/golden/xtc/aa_membrane_peptide.xtc, decoded by the product's xtc reader in the test
    ow = np.array([i for i, (r, a) in enumerate(zip(resn, atn)) if r == "SOL" and a == "OW"], np.uint32)
    hw1 = np.array([i for i, (r, a) in enumerate(zip(resn, atn)) if r == "SOL" and a == "HW1"], np.uint32)
    hw2 = np.array([i for i, (r, a) in enumerate(zip(resn, atn)) if r == "SOL" and a == "HW2"], np.uint32)
    assert len(ow) == len(hw1) == len(hw2) == 5091 and np.all(hw1 == ow + 1) and np.all(hw2 == ow + 2)
    np.savez_compressed(os.path.join(OUT, "aa_membrane_water.npz"), OW=ow, HW1=hw1, HW2=hw2, n_atoms=np.array([len(resn)], np.int64))

    # ---- cfg3: triclinic / dodecahedron / octahedron (50 atoms x 11 frames)
    tri = {}
    for name in ("triclinic", "dodecahedron", "octahedron"):
        gx, gb, _, _ = read_gro(os.path.join(TF, name + ".gro"))
        f, b, t, _ = xdr.read_xtc(os.path.join(TF, name + "_trajectory.xtc"))
        assert f.shape == (11, 50, 3)
        tri[name + "_gro_xyz"], tri[name + "_gro_box"], tri[name + "_frames"], tri[name + "_boxes"] = gx, gb, f, b
    np.savez_compressed(os.path.join(OUT, "triclinic.npz"), **tri)

    # ---- make_group_whole / make_molecules_whole goldens (modifying.rs:1108-1153)
    cxyz, cbox, cbonds = read_pdb(os.path.join(TF, "conect.pdb"))
    wg, wgbox, _, _ = read_gro(os.path.join(TF, "whole_group_expected.gro"))
    wm, wmbox, _, _ = read_gro(os.path.join(TF, "whole_molecules_expected.gro"))
    assert cxyz.shape == wg.shape == wm.shape == (50, 3), (cxyz.shape, wg.shape, wm.shape)
    np.savez_compressed(os.path.join(OUT, "conect.npz"), xyz=cxyz, box=cbox, bonds=cbonds,
                        translate=np.array([3.5, 4.5, -3.0], np.float32), whole_group=wg, whole_molecules=wm)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    sys.exit(main())
