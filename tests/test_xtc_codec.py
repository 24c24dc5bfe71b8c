"""The xtc codec (include/groan_xtc.h; csrc/xtc_codec.hpp, kernels_xtc.cuh) against the reference's own vendored C codec
(external/xdrfile, compiled from where it lies into oracle/_ref/libxdrfile.so) on the reference's xtc fixtures
(tests/golden/xtc/, copied by oracle/copy_xtc_fixtures.py) and on synthetic frames.  Everything here is bit-exact: floats
by bit pattern, files byte for byte.  The host decoder / encoder tests need no GPU; the device decoder tests are `gpu`."""
import os

import numpy as np
import pytest

from oracle import xdrfile_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
XTC = os.path.join(ROOT, "tests", "golden", "xtc")
FIXTURES = ["short_trajectory_protein", "aa_membrane_peptide", "short_trajectory", "triclinic_trajectory",
            "dodecahedron_trajectory", "octahedron_trajectory"]
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libxdrfile.so not built (make -C oracle ref)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _open(name):
    from groan_rs_b200 import xtc
    return xtc.XtcFile.open(os.path.join(XTC, name + ".xtc"))


@needs_ref
@pytest.mark.parametrize("name", FIXTURES + ["short_trajectory_fit"])
def test_host_decoder_is_read_xtc_bit_for_bit(name):
    want = ref.read_xtc(os.path.join(XTC, name + ".xtc"))
    x = _open(name)
    assert (x.n_frames, x.n_atoms) == want["xyz"].shape[:2]
    for threads in (1, 5):
        got = x.decode(n_threads=threads)
        assert np.array_equal(bits(got["xyz"]), bits(want["xyz"]))
        assert np.array_equal(bits(got["box"]), bits(want["box"]))
        assert np.array_equal(got["step"], want["step"]) and np.array_equal(bits(got["time"]), bits(want["time"]))
        assert np.array_equal(bits(got["precision"]), bits(want["precision"]))
    # the integer lattice rebuilds the same floats with the reader's expression (xdrfile.c:844,915-917)
    q = x.decode(want="q32")
    inv = (np.float64(1.0) / want["precision"].astype(np.float64)).astype(np.float32)
    assert np.array_equal(bits(q["q32"].astype(np.float32) * inv[:, None, None]), bits(want["xyz"]))
    q16 = x.decode(want="q16")
    assert np.array_equal(q16["q16"].astype(np.int32) + q16["origin"][:, None, :], q["q32"])
    # a window of frames
    if x.n_frames > 4:
        part = x.decode(first=2, count=3)
        assert np.array_equal(bits(part["xyz"]), bits(want["xyz"][2:5]))


@needs_ref
@pytest.mark.parametrize("name", ["aa_membrane_peptide", "short_trajectory", "triclinic_trajectory"])
def test_partial_frames_like_group_xtc_reader(name):
    """molly_xtc.rs:404-470: only the atoms of a group are produced, decoding stops at the last of them"""
    want = ref.read_xtc(os.path.join(XTC, name + ".xtc"))["xyz"]
    x = _open(name)
    rng = np.random.default_rng(3)
    n = x.n_atoms
    for atoms in (np.arange(0, min(n, 61)), np.sort(rng.choice(n, max(1, n // 7), replace=False)), np.array([n - 1]), np.array([0]),
                  np.arange(n)):
        got = x.decode(atoms=atoms)["xyz"]
        assert np.array_equal(bits(got), bits(want[:, atoms]))
    from groan_rs_b200 import xtc
    with pytest.raises(xtc.XtcError):
        x.decode(atoms=[3, 2])
    with pytest.raises(xtc.XtcError):
        x.decode(atoms=[n])


@needs_ref
@pytest.mark.parametrize("name", FIXTURES + ["short_trajectory_fit"])
def test_encoder_reproduces_the_fixture_files(name):
    """decode -> encode gives back the file byte for byte (the fixtures were written by the same algorithm), from the floats
    and from the lattice integers"""
    from groan_rs_b200 import xtc
    raw = np.fromfile(os.path.join(XTC, name + ".xtc"), dtype=np.uint8)
    x = _open(name)
    d = x.decode()
    prec = float(d["precision"][0])
    again = xtc.encode(xyz=d["xyz"], boxes=d["box"], step=d["step"], time=d["time"], precision=prec, n_threads=3)
    assert np.array_equal(again, raw)
    q = x.decode(want="q32")["q32"]
    again = xtc.encode(q=q, boxes=d["box"], step=d["step"], time=d["time"], precision=prec)
    assert np.array_equal(again, raw)


def _synthetic(kind, n, F, seed):
    rng = np.random.default_rng(seed)
    if kind == "water":  # triplets of atoms 0.1 nm apart on a jittered grid: long runs of small atoms, changing run lengths
        mol = rng.uniform(0.0, 6.0, size=(F, (n + 2) // 3, 1, 3))
        xyz = (mol + rng.normal(0.0, 0.06, size=(F, (n + 2) // 3, 3, 3))).reshape(F, -1, 3)[:, :n]
    elif kind == "gas":  # no spatial order at all: every atom is a large atom
        xyz = rng.uniform(-3.0, 40.0, size=(F, n, 3))
    elif kind == "huge":  # spans more than 2^24 lattice steps: the three-field form of the large atoms (bitsize == 0)
        xyz = rng.uniform(-9000.0, 9000.0, size=(F, n, 3))
        xyz[:, 1::50] = xyz[:, 0::50][:, :xyz[:, 1::50].shape[1]] + 0.05  # some close neighbours (the reference's encoder reads past
        #                                                                   its radix table when there are none at all)
    elif kind == "chain":  # a random walk with small steps: smallidx wanders up and down
        xyz = np.cumsum(rng.normal(0.0, 0.004, size=(F, n, 3)) * rng.choice([1, 1, 1, 30], size=(F, n, 1)), axis=1) + 5.0
    else:
        raise KeyError(kind)
    box = np.tile(np.array([[7.0, 0, 0, 0, 7.5, 0, 1.0, -2.0, 8.0]], np.float32), (F, 1))
    return xyz.astype(np.float32), box, np.arange(F, dtype=np.int32) * 500, np.arange(F, dtype=np.float32) * 2.5


@needs_ref
@pytest.mark.parametrize("kind,n,prec", [("water", 3000, 1000.0), ("water", 301, 100.0), ("gas", 2500, 1000.0), ("huge", 400, 1000.0),
                                         ("chain", 5000, 1000.0), ("chain", 777, 10000.0), ("gas", 10, 1000.0), ("gas", 9, 1000.0),
                                         ("gas", 1, 1000.0)])
def test_encoder_and_decoder_against_write_xtc(kind, n, prec, tmp_path):
    from groan_rs_b200 import xtc
    xyz, box, step, time = _synthetic(kind, n, 4, 11)
    want = ref.write_xtc(str(tmp_path / "ref.xtc"), xyz, box, step, time, prec)
    got = xtc.encode(xyz=xyz, boxes=box, step=step, time=time, precision=prec)
    assert np.array_equal(got, want), (kind, n)
    back = xtc.XtcFile(got)
    assert back.n_frames == 4 and back.n_atoms == n
    rd = ref.read_xtc(str(tmp_path / "ref.xtc"))
    assert np.array_equal(bits(back.decode()["xyz"]), bits(rd["xyz"]))
    if n > 9:
        sel = np.arange(1, n, 3)
        assert np.array_equal(bits(back.decode(atoms=sel)["xyz"]), bits(rd["xyz"][:, sel]))
    else:
        with pytest.raises(xtc.XtcError):
            back.decode(want="q32")  # plain floats: no lattice


def test_damaged_input_is_reported():
    from groan_rs_b200 import xtc
    raw = np.fromfile(os.path.join(XTC, "triclinic_trajectory.xtc"), dtype=np.uint8)
    with pytest.raises(xtc.XtcError) as e:
        xtc.XtcFile(raw[:-5])
    assert e.value.status == 3  # GROAN_XTC_ETRUNC
    bad = raw.copy()
    bad[3] ^= 0x40
    with pytest.raises(xtc.XtcError) as e:
        xtc.XtcFile(bad)
    assert e.value.status == 2  # GROAN_XTC_EMAGIC
    assert xtc.XtcFile(raw[:0]).n_frames == 0
    ok = xtc.XtcFile(raw, max_frames=3)
    assert ok.n_frames == 3


# ---------------------------------------------------------------------------------------------- device decoder
@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("name", ["short_trajectory_protein", "aa_membrane_peptide", "short_trajectory", "triclinic_trajectory",
                                  "dodecahedron_trajectory", "octahedron_trajectory"])
def test_gpu_decoder_is_read_xtc_bit_for_bit(name):
    import groan_rs_b200 as g
    want = ref.read_xtc(os.path.join(XTC, name + ".xtc"))
    x = g.xtc.XtcFile.open(os.path.join(XTC, name + ".xtc"), pinned=True)
    s = g.System(x.n_atoms, max_frames=x.n_frames, triclinic=True)
    meta = s.set_frames_xtc(x)
    assert s.xtc_bad_frames() == 0
    assert np.array_equal(bits(s.get_frames()), bits(want["xyz"]))
    assert np.array_equal(meta["step"], want["step"]) and np.array_equal(bits(meta["time"]), bits(want["time"]))
    assert np.array_equal(bits(s._boxes), bits(want["box"]))
    # a window, from pageable memory
    if x.n_frames > 4:
        y = g.xtc.XtcFile.open(os.path.join(XTC, name + ".xtc"))
        s.set_frames_xtc(y, first=3, count=2)
        assert np.array_equal(bits(s.get_frames()), bits(want["xyz"][3:5]))
    s.close()


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kind,n,prec,F", [("water", 30000, 1000.0, 6), ("gas", 25000, 1000.0, 6), ("huge", 4000, 1000.0, 6),
                                           ("chain", 50000, 1000.0, 6), ("chain", 7777, 10000.0, 6), ("water", 10, 100.0, 6),
                                           ("water", 5000, 1000.0, 600), ("gas", 4100, 1000.0, 3)])
def test_gpu_decoder_synthetic_streams(kind, n, prec, F, tmp_path):
    """one warp per frame (many frames, or small ones) and one CTA per frame (few large frames)"""
    import groan_rs_b200 as g
    xyz, box, step, time = _synthetic(kind, n, F, 5)
    box[:, [3, 6, 7]] = 0
    raw = ref.write_xtc(str(tmp_path / "ref.xtc"), xyz, box, step, time, prec)
    want = ref.read_xtc(str(tmp_path / "ref.xtc"))["xyz"]
    x = g.xtc.XtcFile(raw)
    s = g.System(n, max_frames=F)
    s.set_frames_xtc(x)
    assert s.xtc_bad_frames() == 0
    assert np.array_equal(bits(s.get_frames()), bits(want))
    # results computed on frames decoded by the device equal those on frames uploaded as floats
    c1 = s.group_get_center("all")
    s.set_frames(want, box)
    assert np.array_equal(bits(c1), bits(s.group_get_center("all")))
    s.close()


@pytest.mark.gpu
def test_gpu_decoder_flags_damaged_streams(tmp_path):
    """a frame whose bit stream is shorter than its atoms need (here: the length field of the third frame halved) is reported,
    the frames before it are intact"""
    import groan_rs_b200 as g
    raw = np.fromfile(os.path.join(XTC, "short_trajectory.xtc"), dtype=np.uint8).copy()
    x = g.xtc.XtcFile(raw)
    good = x.decode(count=2)["xyz"]
    lo = int(x.offsets[2])
    nbytes = int.from_bytes(raw[lo + 88:lo + 92].tobytes(), "big")
    raw[lo + 88:lo + 92] = np.frombuffer((nbytes // 2 // 4 * 4).to_bytes(4, "big"), dtype=np.uint8)
    cut = g.xtc.XtcFile(raw, max_frames=3)
    assert cut.n_frames == 3
    s = g.System(x.n_atoms, max_frames=3)
    s.set_frames_xtc(cut)
    assert s.xtc_bad_frames() == 1
    assert np.array_equal(bits(s.get_frames()[:2]), bits(good))
    s.close()


@pytest.mark.gpu
@needs_ref
def test_partial_frames_on_the_device(tmp_path):
    """groan_gpu_push_group_frames: only the group's atoms are uploaded; results equal those of the full frames, the other
    atoms keep their previous values"""
    import groan_rs_b200 as g
    path = os.path.join(XTC, "aa_membrane_peptide.xtc")
    x = g.xtc.XtcFile.open(path)
    full = x.decode()
    n = x.n_atoms
    atoms = np.sort(np.random.default_rng(2).choice(n, 500, replace=False)).astype(np.uint32)
    part = x.decode(atoms=atoms)
    s = g.System(n, max_frames=x.n_frames)
    s.group_create_from_indices("G", atoms)
    s.set_frames(full["xyz"], full["box"])
    c_full = s.group_get_center("G")
    s.set_frames(np.zeros_like(full["xyz"]), full["box"])   # both device slots now hold known values
    s.set_frames(np.full_like(full["xyz"], 7.0), full["box"])
    s.set_group_frames(part["xyz"], atoms, part["box"])
    assert np.array_equal(bits(s.group_get_center("G")), bits(c_full))
    back = s.get_frames()
    assert np.array_equal(bits(back[:, atoms]), bits(full["xyz"][:, atoms]))
    rest = np.setdiff1d(np.arange(n), atoms)
    assert np.all(back[:, rest] == 0.0)  # stale values of the slot (two batches ago), untouched
    s.close()


@pytest.mark.gpu
@needs_ref
def test_fitted_trajectory_is_written_like_the_reference(tmp_path):
    """System.write_xtc: the batch quantised on the device + the host encoder == write_xtc of the same floats"""
    import groan_rs_b200 as g
    x = g.xtc.XtcFile.open(os.path.join(XTC, "short_trajectory.xtc"))
    d = x.decode()
    s = g.System(x.n_atoms, max_frames=x.n_frames)
    s.set_frames(d["xyz"], d["box"])
    s.atoms_translate([0.123, -4.5, 2.0])
    moved = s.get_frames()
    want = ref.write_xtc(str(tmp_path / "ref.xtc"), moved, d["box"], d["step"], d["time"], 100.0)
    got = s.write_xtc(precision=100.0, step=d["step"], time=d["time"])
    assert np.array_equal(got, want)
    s.close()
