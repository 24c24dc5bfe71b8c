"""GPU parity tests: the CUDA path, called through the C ABI (groan_rs_b200.System -> libgroan_gpu.so), against
the CPU oracle and the reference's own golden numbers (file:line cited per test).

Tolerances (BASELINE.json north_star): atom indices and wrap image shifts bit-exact; distances and centres
within 1e-5 nm; RMSD within 1e-4 nm.  All-pairs distances and wrapped positions are additionally required to
be BIT-identical to the ref32 oracle, because the per-pair / per-atom arithmetic is the reference's, f32, no FMA.
"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL_CENTER = 1e-5
TOL_DIST = 1e-5
TOL_RMSD = 1e-4
DIMS = ["None", "X", "Y", "Z", "XY", "XZ", "YZ", "XYZ"]


def _sys(n_atoms, **kw):
    import groan_rs_b200 as g
    return g.System(n_atoms, **kw)


def _dim(name):
    import groan_rs_b200 as g
    return g.Dimension.None_ if name == "None" else g.Dimension[name]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ------------------------------------------------------------------ distances: vector3d.rs:1040-1207
@pytest.mark.parametrize("dim,exp", [("X", 1.5), ("Y", -0.2), ("Z", -1.8), ("XY", 1.51327), ("XZ", 2.34307),
                                     ("YZ", 1.81108), ("XYZ", 2.351595), ("None", 0.0)])
def test_distance_kat(dim, exp):
    s = _sys(2)
    s.group_create_from_indices("a", [0])
    s.group_create_from_indices("b", [1])
    s.set_frames(np.array([[1.0, 3.9, 2.6], [3.5, 0.1, 0.4]], np.float32), [4.0, 4.0, 4.0])
    d = s.group_all_distances("a", "b", _dim(dim))
    assert d.shape == (1, 1, 1)
    assert abs(d[0, 0, 0] - exp) <= 1e-5
    sign = -1.0 if dim in ("X", "Y", "Z") else 1.0
    assert abs(s.group_all_distances("b", "a", _dim(dim))[0, 0, 0] - sign * exp) <= 1e-5
    # out-of-box points, vector3d.rs:1148-1207
    s.set_frames(np.array([[-1.0, 4.5, 2.3], [3.5, -0.5, 4.2]], np.float32), [4.0, 4.0, 4.0])
    for dn, e in (("X", -0.5), ("Y", 1.0), ("Z", -1.9)):
        assert abs(s.group_all_distances("a", "b", _dim(dn))[0, 0, 0] - e) <= 1e-6


def test_min_image_ties_and_far_images():
    """strict comparisons (vector3d.rs:583-589): d == +L/2 stays +L/2; many box lengths away still folds"""
    s = _sys(4)
    s.group_create_from_indices("a", [0, 2])
    s.group_create_from_indices("b", [1, 3])
    xyz = np.array([[3.0, 0, 0], [1.0, 0, 0], [-41.5, 77.0, 1.0], [2.0, -35.25, 9.5]], np.float32)
    s.set_frames(xyz, [4.0, 4.0, 4.0])
    for dn in DIMS:
        got = s.group_all_distances("a", "b", _dim(dn))[0]
        exp = orc.all_distances(xyz, [0, 2], [1, 3], dn, [4.0, 4.0, 4.0])
        assert np.array_equal(bits(got), bits(exp)), dn
    assert s.group_all_distances("a", "b", _dim("X"))[0, 0, 0] == 2.0


# ------------------------------------------------------------------ centres: analysis.rs:488-629, 846-988
def test_center_artificial_kats():
    s = _sys(5, masses=[10.3, 5.4, 3.8, 10.1, 7.6])
    s.group_create_from_indices("g", range(5))
    five = np.array([[3.3, 0.3, 2.5], [4.3, 1.2, 9.8], [3.2, 5.6, 0.5], [0.2, 9.0, 6.6], [8.7, 5.0, 2.4]], np.float32)
    out = np.array([[3.3, 10.3, 2.5], [4.3, 1.2, -0.2], [13.2, 15.6, 0.5], [10.2, -1.0, 6.6], [-1.3, 5.0, 2.4]], np.float32)
    s2 = _sys(5, masses=[10.3, 5.4, 3.8, 10.1, 7.6], max_frames=2)
    s2.group_create_from_indices("g", range(5))
    s2.set_frames(np.stack([five, out]), [10.0, 10.0, 10.0])
    c = s2.group_estimate_center("g")
    for f in range(2):
        assert np.allclose(c[f], [2.634386, 9.775156, 1.1748], atol=1e-4)  # the reference's own epsilon
    assert np.allclose(c[0], orc.estimate_center(five, range(5), [10.0] * 3), atol=TOL_CENTER)
    cm = s2.group_estimate_com("g")
    assert np.allclose(cm[0], [1.9526, 9.7567, 1.8812], atol=1e-4)  # analysis.rs:930-988
    assert np.allclose(cm[0], orc.estimate_center(five, range(5), [10.0] * 3, mass=s2.masses), atol=TOL_CENTER)

    p = _sys(2, masses=[12.8, 0.4])
    p.group_create_from_indices("g", [0, 1])
    pbc = np.array([[4.5, 3.2, 1.7], [9.8, 9.5, 3.0]], np.float32)
    p.set_frames(pbc, [10.0, 10.0, 10.0])
    assert np.allclose(p.group_estimate_center("g")[0], [2.15, 1.35, 2.35], atol=TOL_CENTER)
    assert np.allclose(p.group_get_center("g")[0], [2.15, 1.35, 2.35], atol=TOL_CENTER)
    assert np.allclose(p.group_get_com("g")[0], [4.3575745, 3.0878792, 1.7393947], atol=TOL_CENTER)


def test_center_real_system(example):
    """analysis.rs:631-647, 749-763: example.gro + index.ndx"""
    xyz, box = example["xyz"], example["box"]
    s = _sys(xyz.shape[0])
    s.group_create_from_indices("Protein", example["Protein"])
    s.group_create_from_indices("Membrane", example["Membrane"])
    s.set_frames(xyz, box.reshape(1, 9))
    for name in ("Protein", "Membrane"):
        idx = example[name]
        assert np.allclose(s.group_get_center_naive(name)[0], orc.get_center_naive(xyz, idx), atol=TOL_CENTER)
        assert np.allclose(s.group_estimate_center(name)[0], orc.estimate_center(xyz, idx, box), atol=TOL_CENTER)
        got, exp = s.group_get_center(name)[0], orc.get_center(xyz, idx, box)
        x64 = orc.get_center_x64(xyz, idx, box)
        assert np.allclose(got, exp, atol=TOL_CENTER), (name, got, exp)
        assert np.allclose(got, x64, atol=TOL_CENTER), (name, got, x64)
    assert np.allclose(s.group_get_center_naive("Membrane")[0], [6.47077, 6.52237, 5.77978], atol=1e-4)
    assert np.allclose(s.group_get_center("Protein")[0], [9.85718, 2.46213, 5.45931], atol=1e-4)


def test_center_trajectory_cfg1(protein):
    """BASELINE config 1: protein.gro + short_trajectory_protein.xtc, per-frame group_get_center / group_get_com"""
    fr, bx, m = protein["frames"], protein["boxes"], protein["mass"]
    s = _sys(61, masses=m, max_frames=11)
    s.group_create_from_indices("Protein", range(61))
    s.set_frames(fr, bx)
    c, com = s.group_get_center("Protein"), s.group_get_com("Protein")
    for f in range(11):
        assert np.allclose(c[f], orc.get_center(fr[f], range(61), bx[f]), atol=TOL_CENTER)
        assert np.allclose(com[f], orc.get_com(fr[f], range(61), m, bx[f]), atol=TOL_CENTER)


def test_group_distance_kats(example):
    """analysis.rs:1269-1354 (Protein-Membrane, all dims) -- values from SURVEY 8c"""
    xyz, box = example["xyz"], example["box"]
    s = _sys(xyz.shape[0])
    s.group_create_from_indices("Protein", example["Protein"])
    s.group_create_from_indices("Membrane", example["Membrane"])
    s.set_frames(xyz, box.reshape(1, 9))
    kat = {"X": 6.3029766, "Y": -5.566175, "Z": -0.32046986, "XYZ": 8.415017}
    for dn in DIMS:
        got = s.group_distance("Protein", "Membrane", _dim(dn))[0]
        exp = orc.group_distance(xyz, example["Protein"], example["Membrane"], dn, box)
        assert abs(got - exp) <= TOL_DIST, (dn, got, exp)
        if dn in kat:
            assert abs(got - kat[dn]) <= 1e-4


# ------------------------------------------------------------------ all-pairs: analysis.rs:1420-1530
def test_atoms_distance_kat(example):
    """analysis.rs:1595-1619: System::atoms_distance on example.gro"""
    xyz, box = example["xyz"], example["box"]
    n = xyz.shape[0]
    s = _sys(n)
    s.set_frames(xyz, box.reshape(1, 9))
    for (i, j), exp in (((0, 1), 0.31040135), ((n - 1, 0), 6.664787), ((n - 1, n - 2), 4.062491)):
        d = s.atoms_distance(i, j, _dim("XYZ"))[0]
        assert abs(d - exp) < 1e-5
        assert bits(d) == bits(orc.distance(xyz[i], xyz[j], "XYZ", box.diagonal()))
    assert s.atoms_distance(5, 5, _dim("XYZ"))[0] == 0.0


def test_all_distances_example_bitexact(example):
    xyz, box = example["xyz"], example["box"]
    s = _sys(xyz.shape[0])
    s.group_create_from_indices("Protein", example["Protein"])
    s.group_create_from_indices("Membrane", example["Membrane"])
    s.set_frames(xyz, box.reshape(1, 9))
    for dn in DIMS:
        got = s.group_all_distances("Protein", "Membrane", _dim(dn))[0]
        exp = orc.all_distances(xyz, example["Protein"], example["Membrane"], dn, box)
        assert np.array_equal(bits(got), bits(exp)), dn
    d = s.group_all_distances("Protein", "Protein", _dim("XYZ"))[0]
    assert np.all(np.diag(d) == 0) and np.array_equal(d, d.T)  # analysis.rs:1420-1470
    assert abs(d.max() - 4.597961) <= 1e-5
    # reduce == the documented consumer of the matrix, with Rust's tie rules
    for dn in DIMS:
        r = s.group_all_distances_reduce("Protein", "Membrane", _dim(dn), cutoff=1.0)
        mn, imn, mx, imx, cnt = orc.all_distances_minmax(xyz, example["Protein"], example["Membrane"], dn, box, cutoff=1.0)
        assert bits(r["min"])[0] == bits(mn) and bits(r["max"])[0] == bits(mx), dn
        assert tuple(r["argmin"][0]) == tuple(imn) and tuple(r["argmax"][0]) == tuple(imx), dn
        assert int(r["count"][0]) == cnt, dn


def test_all_distances_cfg2_trajectory(aa_pep):
    """BASELINE config 2: aa_membrane_peptide, peptide -> membrane phosphates, 21 frames, orthogonal box"""
    fr, bx = aa_pep["frames"], aa_pep["boxes"]
    pep, pho = aa_pep["Peptide"], aa_pep["Phosphates"]
    s = _sys(fr.shape[1], max_frames=fr.shape[0])
    s.group_create_from_indices("Peptide", pep)
    s.group_create_from_indices("P", pho)
    s.set_frames(fr, bx)
    got = s.group_all_distances("Peptide", "P", _dim("XYZ"))
    assert got.shape == (21, 363, 128)
    red = s.group_all_distances_reduce("Peptide", "P", _dim("XYZ"), cutoff=0.8)
    gd = s.group_distance("Peptide", "P", _dim("Z"))
    for f in range(fr.shape[0]):
        exp = orc.all_distances(fr[f], pep, pho, "XYZ", bx[f])
        assert np.array_equal(bits(got[f]), bits(exp)), f
        mn, imn, mx, imx, cnt = orc.all_distances_minmax(fr[f], pep, pho, "XYZ", bx[f], cutoff=0.8)
        assert bits(red["min"][f]) == bits(mn) and bits(red["max"][f]) == bits(mx)
        assert tuple(red["argmin"][f]) == tuple(imn) and tuple(red["argmax"][f]) == tuple(imx)
        assert int(red["count"][f]) == cnt
        assert abs(gd[f] - orc.group_distance(fr[f], pep, pho, "Z", bx[f])) <= TOL_DIST


def test_all_distances_ragged_and_empty():
    rng = np.random.default_rng(5)
    n = 1037
    xyz = (rng.random((3, n, 3), dtype=np.float32) * 12 - 1).astype(np.float32)
    L = np.array([[9.5, 10.25, 8.75]] * 3, np.float32)
    s = _sys(n, max_frames=3)
    a = np.arange(3, 3 + 77)            # contiguous, odd size
    b = np.sort(rng.choice(n, 333, replace=False))
    s.group_create_from_indices("a", a)
    s.group_create_from_indices("b", b)
    s.group_create_from_indices("e", [])
    s.set_frames(xyz, L)
    for dn in ("XYZ", "XZ", "Y"):
        got = s.group_all_distances("a", "b", _dim(dn))
        for f in range(3):
            assert np.array_equal(bits(got[f]), bits(orc.all_distances(xyz[f], a, b, dn, L[f]))), (dn, f)
    assert s.group_all_distances("a", "e", _dim("XYZ")).shape == (3, 77, 0)  # empty matrix, not an error
    assert s.group_all_distances("e", "b", _dim("XYZ")).shape == (3, 0, 333)


# ------------------------------------------------------------------ wrap / translate
def test_wrap_kat_and_shifts(example):
    """vector3d.rs:1017-1037 and modifying.rs:688-734"""
    s = _sys(3)
    pts = np.array([[-1.0, 1.5, 3.0], [2.0, 2.2, -0.3], [-54.2, 77.8, 124.5]], np.float32)
    s.set_frames(pts, [2.0, 2.0, 2.0])
    sh = s.atoms_wrap(shifts=True)
    w = s.get_frames()[0]
    exp, esh = orc.wrap(pts, range(3), [2.0, 2.0, 2.0])
    assert np.array_equal(bits(w), bits(exp)) and np.array_equal(sh[0], esh)
    assert w[1, 0] == 2.0  # x == L stays L (strict '>')
    assert np.allclose(w, [[1, 1.5, 1], [2.0, 0.2, 1.7], [1.8, 1.8, 0.5]], atol=1e-5)

    xyz, box = example["xyz"].copy(), example["box"]
    L = np.array([box[0, 0], box[1, 1], box[2, 2]], np.float32)
    rng = np.random.default_rng(11)
    k = rng.integers(-3, 4, size=xyz.shape).astype(np.float32)
    moved = (xyz + k * L).astype(np.float32)
    t = _sys(xyz.shape[0])
    t.group_create_from_indices("Membrane", example["Membrane"])
    t.set_frames(moved, box.reshape(1, 9))
    sh = t.group_wrap("Membrane", shifts=True)
    exp, esh = orc.wrap(moved, example["Membrane"], box)
    got = t.get_frames()[0]
    assert np.array_equal(bits(got), bits(exp))          # untouched atoms included
    assert np.array_equal(sh[0], esh)
    assert np.allclose(got[example["Membrane"]], xyz[example["Membrane"]], atol=1e-4)
    # idempotence
    sh2 = t.group_wrap("Membrane", shifts=True)
    assert not sh2.any() and np.array_equal(bits(t.get_frames()[0]), bits(got))


def test_translate(example):
    """modifying.rs:504-524, atom.rs:1335-1361"""
    xyz, box = example["xyz"], example["box"]
    s = _sys(xyz.shape[0])
    s.group_create_from_indices("Protein", example["Protein"])
    s.set_frames(xyz, box.reshape(1, 9))
    tv = [3.5, -1.1, 5.4]
    sh = s.atoms_translate(tv, shifts=True)
    exp, esh = orc.translate(xyz, range(xyz.shape[0]), tv, box)
    assert np.array_equal(bits(s.get_frames()[0]), bits(exp)) and np.array_equal(sh[0], esh)
    s.group_translate("Protein", [-20.0, 0.25, 31.0])
    exp2, _ = orc.translate(exp, example["Protein"], [-20.0, 0.25, 31.0], box)
    assert np.array_equal(bits(s.get_frames()[0]), bits(exp2))


# ------------------------------------------------------------------ RMSD: rmsd.rs:796-866, 952-1073
GOLDEN_RMSD = [0.23669721, 0.2634763, 0.26021627, 0.21364464, 0.22166993, 0.19383307, 0.26422343, 0.27013618, 0.26398134,
               0.23475659, 0.24208021]


def test_rmsd_cfg1(protein):
    """BASELINE config 1: calc_rmsd of Protein vs the gro frame; golden vector rmsd.rs:811-814"""
    m = protein["mass"]
    ref = _sys(61, masses=m)
    ref.group_create_from_indices("Protein", range(61))
    ref.set_frames(protein["xyz"], protein["box"].reshape(1, 9))
    s = _sys(61, masses=m, max_frames=11)
    s.group_create_from_indices("Protein", range(61))
    s.set_frames(protein["frames"], protein["boxes"])
    rot = np.empty((11, 9), np.float32)
    got = s.calc_rmsd(ref, "Protein", rot=rot)
    for f in range(11):
        e, r = orc.calc_rmsd(protein["xyz"], range(61), protein["box"], m, protein["frames"][f], range(61), protein["boxes"][f])
        assert abs(got[f] - e) <= TOL_RMSD, (f, got[f], e)
        assert np.allclose(rot[f].reshape(3, 3), r, atol=1e-4)
    # the reference's own golden vector (rmsd.rs:811-814); the oracle reproduces it to 1e-6 on these inputs
    assert np.allclose(got, GOLDEN_RMSD, atol=TOL_RMSD)


def test_rmsd_and_fit_short_trajectory(example, short_traj):
    """rmsd.rs:952-1073: fit every frame of short_trajectory.xtc to example.gro; golden short_trajectory_fit.xtc"""
    xyz, box = example["xyz"], example["box"]
    n = xyz.shape[0]
    prot = example["Protein"]
    masses = np.full(n, np.nan, np.float32)
    masses[prot] = short_traj["protein_mass"]
    ref = _sys(n, masses=masses)
    ref.group_create_from_indices("Protein", prot)
    ref.set_frames(xyz, box.reshape(1, 9))
    fr, bx = short_traj["frames"], short_traj["boxes"]
    s = _sys(n, masses=masses, max_frames=fr.shape[0])
    s.group_create_from_indices("Protein", prot)
    s.set_frames(fr, bx)
    rm = s.calc_rmsd(ref, "Protein")
    assert np.allclose(rm, GOLDEN_RMSD, atol=TOL_RMSD)
    rm2 = s.calc_rmsd_and_fit(ref, "Protein")
    assert np.array_equal(rm, rm2)
    fitted = s.get_frames()
    for f in range(fr.shape[0]):
        e, ef = orc.calc_rmsd_and_fit(xyz, prot, box, short_traj["protein_mass"], fr[f], prot, bx[f])
        assert abs(rm[f] - e) <= TOL_RMSD
        assert np.abs(fitted[f] - ef).max() <= 2e-4, f
        # golden fitted trajectory, quantised to 0.01 nm (half quantum + 7e-5 with the gro reference)
        assert np.abs(fitted[f] - short_traj["fit"][f]).max() <= 0.0053, f  # the ref32 oracle itself: 0.00524 (frame 7)
    # the lattice points the xtc writer would store (groan_gpu_get_frames_quantized: xdrfile.c:1018-1031) against the golden
    # file's.  Byte-identity is out of reach without nalgebra's own f32 SVD (an un-vendored dependency): the reference's
    # rotation differs from the exact one by ~3e-5, which moves 0.2-0.8 % of the coordinates across a rounding boundary of
    # the 0.01 nm lattice -- always by exactly one step (the CPU restatement shows the same: DESIGN.md section 10).
    q = s.get_frames_quantized(100.0)
    gq = np.rint(short_traj["fit"] * 100.0).astype(np.int64)
    d = q.astype(np.int64) - gq
    assert np.abs(d).max() <= 1
    assert (d != 0).mean() <= 0.01, (d != 0).mean()


def test_rmsd_identity_and_broken_reference(protein):
    """rmsd.rs:844-866: a PBC-broken copy of the same structure has RMSD 0"""
    m = protein["mass"]
    L = np.array([protein["box"][0, 0], protein["box"][1, 1], protein["box"][2, 2]], np.float32)
    ref = _sys(61, masses=m)
    ref.group_create_from_indices("Protein", range(61))
    ref.set_frames(protein["xyz"], protein["box"].reshape(1, 9))
    broken = protein["xyz"].copy()
    broken[::2] += L
    broken[1::3] -= L
    s = _sys(61, masses=m, max_frames=2)
    s.group_create_from_indices("Protein", range(61))
    s.set_frames(np.stack([protein["xyz"], broken]), np.stack([protein["box"]] * 2))
    got = s.calc_rmsd(ref, "Protein")
    assert np.all(np.abs(got) <= 1e-4), got


def _check_reduce_against_matrix(s, ga, gb, dn, cutoff):
    """group_all_distances_reduce == the reference's scan of the materialised matrix (analysis.rs:390-399), bit for bit"""
    mat = s.group_all_distances(ga, gb, _dim(dn))
    red = s.group_all_distances_reduce(ga, gb, _dim(dn), cutoff=cutoff)
    F, n1, n2 = mat.shape
    for f in range(F):
        m = mat[f]
        kmin = int(np.argmin(m))                      # first minimum in row-major order
        kmax = m.size - 1 - int(np.argmax(m[::-1, ::-1]))  # last maximum in row-major order
        assert bits(red["min"][f]) == bits(m.flat[kmin]) and tuple(red["argmin"][f]) == (kmin // n2, kmin % n2), (dn, f)
        assert bits(red["max"][f]) == bits(m.flat[kmax]) and tuple(red["argmax"][f]) == (kmax // n2, kmax % n2), (dn, f)
        assert int(red["count"][f]) == int(np.count_nonzero(m < np.float32(cutoff))), (dn, f)


def test_triclinic_reduce_fast_path_synthetic():
    """the 27-image d^2 of the fused reduction against the materialised triclinic matrix (itself bit-exact against the
    oracle, test_triclinic_cfg3) on a dodecahedron-like box with atoms up to three box vectors outside the cell"""
    n, F = 6000, 2
    box = np.array([7.0, 0, 0, 0, 7.0, 0, 3.5, 3.5, 4.95], np.float32)
    s = _sys(n, max_frames=F, triclinic=True)
    rng = np.random.default_rng(21)
    fr = rng.uniform(-14.0, 21.0, (F, n, 3)).astype(np.float32)
    fr[:, 100] = fr[:, 3]  # a zero distance, and a tie with it
    fr[:, 4000] = fr[:, 3]
    s.set_frames(fr, box)
    s.group_create_from_indices("A", range(0, 70))
    s.group_create_from_indices("B", range(64, n))
    for dn in ("XYZ", "XZ", "XY"):
        _check_reduce_against_matrix(s, "A", "B", dn, cutoff=1.2)


# ------------------------------------------------------------------ triclinic EXTENSION (config 3; unpinned by the reference)
@pytest.mark.parametrize("name", ["triclinic", "dodecahedron", "octahedron"])
def test_triclinic_cfg3(tric, name):
    import groan_rs_b200 as g
    fr, bx = tric[name + "_frames"], tric[name + "_boxes"]
    n = fr.shape[1]
    # without the extension flag the behaviour is the reference's: NotOrthogonal (simbox.rs:230-236)
    plain = _sys(n, max_frames=fr.shape[0])
    plain.set_frames(fr, bx)
    with pytest.raises(g.GroanError) as ei:
        plain.atoms_wrap()
    assert "NotOrthogonal" in ei.value.variant
    with pytest.raises(g.GroanError):
        plain.group_all_distances("all", "all", g.Dimension.XYZ)

    s = _sys(n, max_frames=fr.shape[0], triclinic=True)
    s.set_frames(fr, bx)
    for dn in ("XYZ", "XY", "Z"):
        got = s.group_all_distances("all", "all", _dim(dn))
        for f in range(fr.shape[0]):
            exp = orc.tric_all_distances(fr[f], range(n), range(n), dn, bx[f])
            assert np.array_equal(bits(got[f]), bits(exp)), (dn, f)
    # the fused reduction (27-image d^2 on the fast path for 2-D / 3-D, reference loops for 1-D) = reduction of that matrix:
    # first minimum, last maximum of the row-major scan, count below the cutoff
    s.group_create_from_indices("A", range(0, 20))
    s.group_create_from_indices("B", range(15, n))
    for dn in ("XYZ", "XY", "YZ", "Z"):
        _check_reduce_against_matrix(s, "A", "B", dn, cutoff=0.9)
    _check_reduce_against_matrix(s, "all", "all", "XYZ", cutoff=0.5)
    # self-pin: f64 brute force over 5^3 images
    got = s.group_all_distances("all", "all", _dim("XYZ"))
    for f in (0, 5, 10):
        for i in range(0, n, 7):
            for j in range(n):
                b = orc.tric_distance_brute64(fr[f, i], fr[f, j], "XYZ", bx[f], 2)
                assert abs(got[f, i, j] - b) <= TOL_DIST
    # wrap after moving atoms by random integer combinations of box vectors: shifts bit-exact vs the ref32 oracle
    rng = np.random.default_rng(3)
    k = rng.integers(-2, 3, size=(fr.shape[0], n, 3)).astype(np.float32)
    moved = fr.copy()
    for f in range(fr.shape[0]):
        B = bx[f].reshape(3, 3)
        moved[f] = (fr[f] + k[f] @ B).astype(np.float32)
    s.set_frames(moved, bx)
    sh = s.atoms_wrap(shifts=True)
    w = s.get_frames()
    for f in range(fr.shape[0]):
        exp, esh = orc.tric_wrap(moved[f], range(n), bx[f])
        assert np.array_equal(sh[f], esh), f
        assert np.array_equal(bits(w[f]), bits(exp)), f
    # centre / RMSD: the reference's behaviour without the flag (with it: test_triclinic_centre_and_rmsd_*)
    with pytest.raises(g.GroanError) as ei:
        plain.group_get_center("all")
    assert "NotOrthogonal" in ei.value.variant


@pytest.mark.parametrize("name", ["triclinic", "dodecahedron", "octahedron"])
def test_triclinic_centre_and_rmsd_fixtures(tric, name):
    """Triclinic EXTENSION of group_estimate_center / group_get_center / group_get_com / calc_rmsd (parity unpinned: the
    reference returns NotOrthogonal).  Definition = oracle.tric_*_x64 (Bai-Breen on fractional coordinates, i.e. the
    orthogonal algorithm in the sheared picture); here on the reference's three 50-atom triclinic trajectories."""
    fr, bx = tric[name + "_frames"], tric[name + "_boxes"]
    F, n = fr.shape[0], fr.shape[1]
    masses = np.random.default_rng(12).uniform(1.0, 30.0, n).astype(np.float32)
    s = _sys(n, masses=masses, max_frames=F, triclinic=True)
    s.set_frames(fr, bx)
    idx = np.arange(3, n - 2)
    s.group_create_from_indices("G", idx)
    ref = _sys(n, masses=masses, triclinic=True)
    ref.set_frames(fr[0], bx[0].reshape(1, 9))
    ref.group_create_from_indices("G", idx)
    est, cen, com = s.group_estimate_center("G"), s.group_get_center("G"), s.group_get_com("G")
    rm = s.calc_rmsd(ref, "G")
    assert abs(float(rm[0])) <= TOL_RMSD  # frame 0 against itself
    for f in range(F):
        assert np.abs(est[f] - orc.tric_estimate_center_x64(fr[f], idx, bx[f])).max() <= 2e-5, (f, est[f])
        assert np.abs(cen[f] - orc.tric_get_center_x64(fr[f], idx, bx[f])).max() <= TOL_CENTER, (f, cen[f])
        assert np.abs(com[f] - orc.tric_get_center_x64(fr[f], idx, bx[f], masses[idx])).max() <= TOL_CENTER, (f, com[f])
        e, _ = orc.tric_calc_rmsd_x64(fr[0], idx, bx[0], masses[idx], fr[f], idx, bx[f])
        assert abs(float(rm[f]) - e) <= TOL_RMSD, (f, rm[f], e)
    # the fit stays orthogonal-only, with or without the flag (fit_structure is not part of the extension)
    import groan_rs_b200 as g
    with pytest.raises(g.GroanError) as ei:
        s.calc_rmsd_and_fit(ref, "G")
    assert "NotOrthogonal" in ei.value.variant


def test_triclinic_centre_and_rmsd_blob_and_orthogonal_identity():
    """Self-pins of the triclinic centre / RMSD extension.
    (ii) a compact blob, rigidly moved and wrapped into a triclinic cell: the centre must be the plain mean of the blob
         unwrapped by brute force (nearest of 5^3 periodic images to its first atom, Cartesian metric, float64) up to a whole
         box vector combination, and the RMSD the noise level it was built with -- neither involves the shear.
    (i)  orthogonal frames of a batch that also holds a triclinic frame go through the sheared code path (the shear of an
         orthogonal box is the identity) and must equal the orthogonal-only run of the same kernels bit for bit."""
    import groan_rs_b200 as g
    n, F = 120_000, 4
    box = np.array([12.0, 0, 0, 3.0, 11.0, 0, 2.5, -3.5, 10.0], np.float32)
    orth = np.array([12.0, 0, 0, 0, 11.0, 0, 0, 0, 10.0], np.float32)
    masses = np.random.default_rng(3).uniform(1.0, 100.0, n).astype(np.float32)
    scale, nscale = 1.2 / 131070.0, 0.03 / 37837.23
    rot = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (F, 1))
    rot[1] = [0, -1, 0, 1, 0, 0, 0, 0, 1]
    rot[2] = [1, 0, 0, 0, 0, -1, 0, 1, 0]
    cen = np.array([[6.0, 5.0, 5.0], [0.4, 10.6, 9.8], [13.5, 3.0, 0.2], [2.0, 2.0, 2.0]], np.float32)
    gen = g.System(n, masses=masses, max_frames=F)
    gen.synth_blob(17, 0, F, scale, nscale, rot, cen, [50.0] * 3, wrap=False)  # unwrapped blobs; the box given here is not used
    raw = gen.get_frames().copy()
    ref_xyz = gen.synth_blob_ref(17, scale, [6.0, 5.0, 5.0])
    gen.close()
    boxes = np.tile(box, (F, 1))
    boxes[3] = orth
    s = g.System(n, masses=masses, max_frames=F, triclinic=True)
    s.set_frames(raw, boxes)
    s.atoms_wrap()  # into the triclinic cell (frame 3: the orthogonal box)
    wrapped = s.get_frames().copy()
    idx = np.arange(5, n - 3)
    s.group_create_from_indices("G", idx)
    ref = g.System(n, masses=masses, triclinic=True)
    ref.set_frames(ref_xyz, box.reshape(1, 9))
    ref.group_create_from_indices("G", idx)
    cen_g, com_g, rm = s.group_get_center("G"), s.group_get_com("G"), s.calc_rmsd(ref, "G")
    imgs = np.array([[i, j, k] for i in range(-2, 3) for j in range(-2, 3) for k in range(-2, 3)], np.float64)
    for f in range(F):
        B = boxes[f].reshape(3, 3).astype(np.float64)
        x = wrapped[f, idx].astype(np.float64)
        shifts = imgs @ B
        d2 = (((x - x[0])[:, None, :] + shifts[None, :, :]) ** 2).sum(axis=2)
        un = x + shifts[np.argmin(d2, axis=1)]
        for got, want in ((cen_g[f], un.mean(axis=0)), (com_g[f], (un * masses[idx, None]).sum(axis=0) / masses[idx].sum())):
            k = np.linalg.solve(B.T, got.astype(np.float64) - want)  # difference in units of the box vectors
            assert np.abs(k - np.rint(k)).max() * 12.0 <= 2e-5, (f, got, want, k)
        assert abs(float(rm[f]) - 0.03 * np.sqrt(3.0)) <= 2e-3, (f, rm[f])  # noise sigma 0.03 nm per axis
        e, _ = orc.tric_calc_rmsd_x64(ref_xyz, idx, box, masses[idx], wrapped[f], idx, boxes[f])
        assert abs(float(rm[f]) - e) <= TOL_RMSD, (f, rm[f], e)
        assert np.abs(cen_g[f] - orc.tric_get_center_x64(wrapped[f], idx, boxes[f])).max() <= TOL_CENTER
    # (i): frame 3 (orthogonal box) inside the mixed batch == the same frame through the orthogonal code, bit for bit.  A large
    # contiguous group runs on the rings in both cases (k_center_quad / k_rmsd_quad with and without the shear) ...
    oref = g.System(n, masses=masses)
    oref.set_frames(ref_xyz, orth.reshape(1, 9))
    oref.group_create_from_indices("G", idx)
    sref = g.System(n, masses=masses, triclinic=True)
    sref.set_frames(ref_xyz, orth.reshape(1, 9))
    sref.group_create_from_indices("G", idx)
    oq = g.System(n, masses=masses, max_frames=F)  # the same batch size: the same grids, hence the same order of the partial sums
    oq.set_frames(wrapped, np.tile(orth, (F, 1)))
    oq.group_create_from_indices("G", idx)
    rm_ring = s.calc_rmsd(sref, "G")
    assert np.array_equal(bits(oq.group_get_center("G")[3]), bits(cen_g[3]))
    assert np.array_equal(bits(oq.group_get_com("G")[3]), bits(com_g[3]))
    assert np.array_equal(bits(oq.calc_rmsd(oref, "G")[3:4]), bits(rm_ring[3:4]))
    oq.close()
    # ... and with GROAN_FLAG_NO_TMA through the gather kernels in both cases; ring and gather agree on every frame
    s.set_flags(g.FLAG_TRICLINIC | g.FLAG_NO_TMA)
    cen_n, com_n, rm_n, rm_mixed = s.group_get_center("G"), s.group_get_com("G"), s.calc_rmsd(ref, "G"), s.calc_rmsd(sref, "G")
    s.set_flags(g.FLAG_TRICLINIC)
    assert np.abs(cen_n - cen_g).max() <= 4e-6 and np.abs(com_n - com_g).max() <= 4e-6 and np.abs(rm_n - rm).max() <= 2e-6
    o = g.System(n, masses=masses, max_frames=F)
    o.set_flags(g.FLAG_NO_TMA)
    o.set_frames(wrapped, np.tile(orth, (F, 1)))
    o.group_create_from_indices("G", idx)
    assert np.array_equal(bits(o.group_get_center("G")[3]), bits(cen_n[3]))
    assert np.array_equal(bits(o.group_get_com("G")[3]), bits(com_n[3]))
    assert np.array_equal(bits(o.calc_rmsd(oref, "G")[3:4]), bits(rm_mixed[3:4]))


def test_triclinic_code_on_orthogonal_box_is_identical(example):
    """self-pin (i): with the flag set, an orthogonal box goes through the orthogonal kernels and matches bit for bit"""
    xyz, box = example["xyz"], example["box"]
    s = _sys(xyz.shape[0], triclinic=True)
    s.group_create_from_indices("Protein", example["Protein"])
    s.set_frames(xyz, box.reshape(1, 9))
    got = s.group_all_distances("Protein", "Protein", _dim("XYZ"))[0]
    assert np.array_equal(bits(got), bits(orc.all_distances(xyz, example["Protein"], example["Protein"], "XYZ", box)))
    assert np.array_equal(bits(got), bits(orc.tric_all_distances(xyz, example["Protein"], example["Protein"], "XYZ", box)))


# ------------------------------------------------------------------ errors: analysis.rs:650-746, rmsd.rs:1076-1183
def test_error_paths(protein):
    import groan_rs_b200 as g
    m = protein["mass"]
    s = _sys(61, masses=m, max_frames=2)
    s.group_create_from_indices("Protein", range(61))
    s.group_create_from_indices("Empty", [])
    with pytest.raises(g.GpuError):  # no frames yet
        s.group_get_center("Protein")
    s.set_frames(protein["frames"][:2], None)
    with pytest.raises(g.GroupError) as ei:
        s.group_get_center("Protein")
    assert "DoesNotExist" in ei.value.variant
    s.set_frames(protein["frames"][:2], protein["boxes"][:2])
    with pytest.raises(g.GroupError) as ei:
        s.group_get_center("Nonexistent")
    assert ei.value.variant == "NotFound"
    with pytest.raises(g.GroupError) as ei:
        s.group_get_center("Empty")
    assert ei.value.variant == "EmptyGroup"
    with pytest.raises(g.GroupError) as ei:
        s.group_distance("Protein", "Empty", g.Dimension.XYZ)
    assert ei.value.variant == "EmptyGroup"
    valid = np.ones((2, 61), np.uint8)
    valid[1, 17] = 0
    s.set_valid(valid)
    with pytest.raises(g.GroupError) as ei:
        s.group_get_center("Protein")
    assert "NoPosition(17)" in ei.value.variant and ei.value.detail == (1, 17)
    with pytest.raises(g.GroupError):
        s.atoms_wrap()
    s.set_valid(None)
    s.group_get_center("Protein")
    # zero box: the reference panics "Box len should not be zero"
    zb = protein["boxes"][:2].copy()
    zb[1, 2, 2] = 0.0
    s.set_frames(protein["frames"][:2], zb)
    with pytest.raises(g.GroanError):
        s.group_get_center("Protein")
    # masses
    nm = _sys(61, max_frames=1)
    nm.group_create_from_indices("Protein", range(61))
    nm.set_frames(protein["frames"][:1], protein["boxes"][:1])
    with pytest.raises(g.GroupError) as ei:
        nm.group_get_com("Protein")
    assert "NoMass" in ei.value.variant
    # RMSD: inconsistent group sizes (rmsd.rs:1110-1130), missing group, missing masses
    ref = _sys(61, masses=m)
    ref.group_create_from_indices("Protein", range(60))
    ref.set_frames(protein["xyz"], protein["box"].reshape(1, 9))
    s.set_frames(protein["frames"][:2], protein["boxes"][:2])
    with pytest.raises(g.RMSDError) as ei:
        s.calc_rmsd(ref, "Protein")
    assert ei.value.variant == "InconsistentGroup" and ei.value.detail == (60, 61)
    with pytest.raises(g.RMSDError) as ei:
        s.calc_rmsd(ref, "Nope")
    assert ei.value.variant == "NonexistentGroup"
    with pytest.raises(g.GpuError):  # capacity
        s.set_frames(protein["frames"][:3], protein["boxes"][:3])


# ------------------------------------------------------------------ synthetic generators == oracle generators
def test_synth_generators_match_oracle():
    n = 5000
    s = _sys(n, max_frames=3)
    lo, span = [-2.15, -2.15, -2.15], [25.8, 25.8, 25.8]
    s.synth_uniform(20261018, 7, 3, lo, span, [21.5, 21.5, 21.5])
    got = s.get_frames()
    for f in range(3):
        assert np.array_equal(bits(got[f]), bits(orc.synth_uniform(n, 20261018, 7 + f, lo, span)))
    rot = np.stack([np.eye(3, dtype=np.float32)] * 3)
    th = 0.3
    rot[1] = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    cen = np.array([[3, 4, 5], [33.5, 0.2, 17], [10, 20, 30]], np.float32)
    L = [34.0, 34.0, 34.0]
    s.synth_blob(99, 100, 3, 2.5e-5, 1e-7, rot, cen, L, wrap=True)
    got = s.get_frames()
    for f in range(3):
        exp = orc.synth_blob_frame(n, 99, 100 + f, 2.5e-5, 1e-7, rot[f], cen[f], L, wrap=True)
        assert np.array_equal(bits(got[f]), bits(exp)), f
    assert np.array_equal(bits(s.synth_blob_ref(99, 2.5e-5, [17, 17, 17])), bits(orc.synth_blob_ref(n, 99, 2.5e-5, [17, 17, 17])))


# ------------------------------------------------------------------ single-pass kernels vs reference-order passes
def _blob_system(n, F, L, seed, scale, nscale, masses, flags=0):
    import groan_rs_b200 as g
    rng = np.random.default_rng(seed)
    rot = np.empty((F, 9), np.float32)
    for k in range(F):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        rot[k] = [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                  2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]
    cen = rng.uniform(0, L, size=(F, 3)).astype(np.float32)
    cen[0] = [0.01, L[1] - 0.01, L[2] / 2]  # straddles two box faces
    s = g.System(n, masses=masses, max_frames=F)
    if flags:
        s.set_flags(flags)
    s.synth_blob(seed, 0, F, scale, nscale, rot, cen, L, wrap=True)
    return s


@pytest.mark.parametrize("n,contig", [(200_003, True), (200_003, False), (1_000_000, True)])
def test_single_pass_matches_exact_and_x64(n, contig):
    """group_get_center / group_get_com / calc_rmsd: the single-pass kernels, the reference-order passes
    (GROAN_FLAG_EXACT_ONLY) and the exact64 oracle agree within the north-star tolerances at a size where the
    reference's sequential f32 sums have already drifted."""
    import groan_rs_b200 as g
    F, L = 4, np.array([20.0, 21.0, 19.0], np.float32)
    rng = np.random.default_rng(1)
    masses = rng.uniform(1.0, 100.0, n).astype(np.float32)
    idx = np.arange(5, n - 3) if contig else np.sort(rng.choice(n, n // 2, replace=False))
    scale, nscale = 3.0 / 131070.0, 0.03 / 37837.23
    fast = _blob_system(n, F, L, 77, scale, nscale, masses)
    exact = _blob_system(n, F, L, 77, scale, nscale, masses, flags=g.FLAG_EXACT_ONLY)
    ref = g.System(n, masses=masses)
    ref_xyz = fast.synth_blob_ref(77, scale, L / 2)
    ref.set_frames(ref_xyz, L)
    for s in (fast, exact, ref):
        s.group_create_from_indices("G", idx)
    frames = fast.get_frames()
    l0 = fast.launch_count()
    cf = fast.group_get_center("G")
    assert fast.fallback_frames() == 0  # compact group, centre away from the box edge: certified by the single pass
    mf = fast.group_get_com("G")
    assert fast.fallback_frames() == 0
    rf = fast.calc_rmsd(ref, "G")
    assert fast.fallback_frames() == 0
    assert fast.launch_count() - l0 >= 3
    # centre + RMSD from one read of the frame == the two separate calls
    for weighted, sep in ((False, cf), (True, mf)):
        c2, r2 = fast.group_center_and_rmsd(ref, "G", weighted=weighted)
        assert fast.fallback_frames() == 0
        assert np.abs(c2 - sep).max() <= 4e-6, np.abs(c2 - sep).max()  # 2 ulp of a 20 nm coordinate
        assert np.abs(r2 - rf).max() <= 2e-6, np.abs(r2 - rf).max()
    ce, me, re_ = exact.group_get_center("G"), exact.group_get_com("G"), exact.calc_rmsd(ref, "G")
    c3, r3 = exact.group_center_and_rmsd(ref, "G")
    assert np.array_equal(bits(c3), bits(ce)) and np.array_equal(bits(r3), bits(re_))
    assert np.abs(cf - ce).max() <= TOL_CENTER and np.abs(mf - me).max() <= TOL_CENTER
    assert np.abs(rf - re_).max() <= 2e-5
    for f in range(F):
        c64 = orc.get_center_x64(frames[f], idx, L)
        m64 = orc.get_center_x64(frames[f], idx, L, mass=masses[idx])
        r64, _ = orc.calc_rmsd_x64(ref_xyz, idx, L, masses[idx], frames[f], idx, L)
        assert np.abs(cf[f] - c64).max() <= TOL_CENTER, (f, cf[f], c64)
        assert np.abs(mf[f] - m64).max() <= TOL_CENTER, (f, mf[f], m64)
        assert abs(rf[f] - r64) <= 2e-5, (f, rf[f], r64)
        assert abs(r64 - np.sqrt(3) * 0.03) < 1e-3  # analytic RMSD of the generator


def test_single_pass_falls_back_when_it_cannot_certify(example):
    """box-spanning group (Membrane spans x and y), c0 on the box edge, and RMSD ~ 0: the single-pass kernels flag
    the frame and the reference-order passes produce the result -- identical to GROAN_FLAG_EXACT_ONLY."""
    import groan_rs_b200 as g
    xyz, box = example["xyz"], example["box"]
    n = xyz.shape[0]
    res = []
    for flags in (0, g.FLAG_EXACT_ONLY):
        s = g.System(n, masses=np.ones(n, np.float32))
        s.set_flags(flags)
        s.group_create_from_indices("Membrane", example["Membrane"])
        s.set_frames(xyz, box.reshape(1, 9))
        res.append((s.group_get_center("Membrane"), s.group_get_com("Membrane")))
    assert np.array_equal(bits(res[0][0]), bits(res[1][0])) and np.array_equal(bits(res[0][1]), bits(res[1][1]))
    assert s.fallback_frames() == 0  # EXACT_ONLY never flags
    # symmetric pair around the box edge: the circular mean sits on the edge, where k is decided by rounding
    two = np.array([[9.9, 5.0, 0.2], [0.1, 5.0, 0.4]], np.float32)
    out = []
    for flags in (0, g.FLAG_EXACT_ONLY):
        s = g.System(2)
        s.set_flags(flags)
        s.group_create_from_indices("g", [0, 1])
        s.set_frames(two, [10.0, 10.0, 10.0])
        out.append(s.group_get_center("g"))
        assert s.fallback_frames() == (1 if flags == 0 else 0)
    assert np.array_equal(bits(out[0]), bits(out[1]))
    assert np.allclose(out[0][0], orc.get_center(two, [0, 1], [10.0] * 3), atol=TOL_CENTER)


@pytest.mark.parametrize("far", [False, True])
def test_all_pairs_fast_paths_bitexact(far):
    """2-D / 3-D all-pairs on an orthogonal box: coordinates up to 10 % outside the box take the packed one-step fold,
    a few atoms many box lengths away force the loop version for their tiles; both must equal the ref32 oracle bit for bit,
    matrix and fused reduction (first minimum / last maximum / count below cutoff)."""
    n, F = 6000, 2
    L = np.array([7.5, 8.25, 6.75], np.float32)
    s = _sys(n, max_frames=F)
    s.synth_uniform(123, 0, F, -0.1 * L, 1.2 * L, L)
    frames = s.get_frames()
    if far:
        frames[0, 17] += np.float32(5) * L
        frames[1, 4100] -= np.float32(9) * L
        frames[1, 4101, 1] += np.float32(1.6) * L[1]
        s.set_frames(frames, L)
    a = np.arange(0, 1900)          # not a multiple of the row tile
    b = np.arange(3000, 3000 + 2603)  # odd size: scalar stores, ragged last thread
    s.group_create_from_indices("A", a)
    s.group_create_from_indices("B", b)
    s.group_create_from_indices("B4", b[:2600])  # multiple of 4: float4 stores
    for dn in ("XYZ", "XY", "XZ", "YZ"):
        got = s.group_all_distances("A", "B", _dim(dn))
        got4 = s.group_all_distances("A", "B4", _dim(dn))
        red = s.group_all_distances_reduce("A", "B", _dim(dn), cutoff=0.35)
        for f in range(F):
            exp = orc.all_distances(frames[f], a, b, dn, L)
            assert np.array_equal(bits(got[f]), bits(exp)), (dn, f)
            assert np.array_equal(bits(got4[f]), bits(exp[:, :2600])), (dn, f)
            mn, imn, mx, imx, cnt = orc.all_distances_minmax(frames[f], a, b, dn, L, cutoff=0.35)
            assert bits(red["min"][f]) == bits(mn) and bits(red["max"][f]) == bits(mx), (dn, f)
            assert tuple(red["argmin"][f]) == tuple(imn) and tuple(red["argmax"][f]) == tuple(imx), (dn, f)
            assert int(red["count"][f]) == cnt, (dn, f)


def test_all_pairs_reduce_ties_and_many_chunks():
    """ties: duplicated atoms give equal distances at several (i, j); the fused reduction must return the FIRST minimum
    and the LAST maximum of the row-major scan even when a CTA sweeps group B in several chunks (64 frames cap the grid)."""
    n, F = 140_000, 64
    L = np.array([9.0, 9.0, 9.0], np.float32)
    s = _sys(n, max_frames=F)
    s.synth_uniform(7, 0, F, [0, 0, 0], L, L)
    frames = s.get_frames()
    # plant ties: copies of A atoms inside B (distance 0 at several positions) and mirrored far pairs
    frames[:, 200] = frames[:, 3]
    frames[:, 139_000] = frames[:, 3]
    frames[:, 70_000] = frames[:, 1]
    s.set_frames(frames, L)
    a = np.arange(0, 6)
    b = np.arange(100, n)
    s.group_create_from_indices("A", a)
    s.group_create_from_indices("B", b)
    red = s.group_all_distances_reduce("A", "B", _dim("XYZ"), cutoff=0.5)
    for f in (0, 31, 63):
        mn, imn, mx, imx, cnt = orc.all_distances_minmax(frames[f], a, b, "XYZ", L, cutoff=0.5)
        assert mn == 0.0 and red["min"][f] == 0.0
        assert tuple(red["argmin"][f]) == tuple(imn), (f, red["argmin"][f], imn)
        assert bits(red["max"][f]) == bits(mx) and tuple(red["argmax"][f]) == tuple(imx)
        assert int(red["count"][f]) == cnt


def test_all_pairs_reduce_device_outputs_and_warp_shared_thresholds():
    """results delivered into device tensors (no host copies) equal the host-delivered ones, on a case where the running
    minimum / maximum of different lanes of a warp keep improving (sorted distances: every step is a record for some lane),
    i.e. the thresholds the warp shares after an update are exercised all the time."""
    import torch
    n, F = 40_000, 3
    L = np.array([30.0, 30.0, 30.0], np.float32)
    s = _sys(n, max_frames=F)
    frames = np.zeros((F, n, 3), np.float32)
    rng = np.random.default_rng(11)
    for f in range(F):
        # group A near the origin, group B on a line of decreasing then increasing distance
        frames[f, :16] = rng.uniform(0.0, 0.2, (16, 3))
        x = np.linspace(14.9, 0.3, n - 16) if f != 1 else np.linspace(0.3, 14.9, n - 16)
        frames[f, 16:, 0] = x
        frames[f, 16:, 1] = rng.uniform(0.0, 0.05, n - 16)
    s.set_frames(frames, L)
    a, b = np.arange(0, 16), np.arange(16, n)
    s.group_create_from_indices("A", a)
    s.group_create_from_indices("B", b)
    host = s.group_all_distances_reduce("A", "B", _dim("XYZ"), cutoff=1.0)
    dev = torch.device("cuda", 0)
    out = {"min": torch.empty(F, dtype=torch.float32, device=dev), "argmin": torch.empty((F, 2), dtype=torch.int32, device=dev),
           "max": torch.empty(F, dtype=torch.float32, device=dev), "argmax": torch.empty((F, 2), dtype=torch.int32, device=dev),
           "count": torch.empty(F, dtype=torch.int64, device=dev)}
    s.group_all_distances_reduce("A", "B", _dim("XYZ"), cutoff=1.0, out=out)
    s.sync()
    for k in host:
        assert np.array_equal(out[k].cpu().numpy().astype(np.int64 if k != "min" and k != "max" else np.float32),
                              host[k].astype(np.int64 if k != "min" and k != "max" else np.float32)), k
    for f in range(F):
        mn, imn, mx, imx, cnt = orc.all_distances_minmax(frames[f], a, b, "XYZ", L, cutoff=1.0)
        assert bits(host["min"][f]) == bits(mn) and tuple(host["argmin"][f]) == tuple(imn), (f, host["argmin"][f], imn)
        assert bits(host["max"][f]) == bits(mx) and tuple(host["argmax"][f]) == tuple(imx), (f, host["argmax"][f], imx)
        assert int(host["count"][f]) == cnt


# ------------------------------------------------------------------ full-size, size-independent properties (BASELINE configs[3], [4])
def test_full_size_properties_cfg4_cfg5():
    """At BASELINE.json's full sizes the oracle is too slow, so the CUDA path is checked through properties:
    translation by box vectors leaves centre (mod L) and RMSD unchanged; a rigidly rotated noise-free copy has RMSD 0 and
    the rotation is recovered; wrap is idempotent and undoes integer box shifts; the fused all-pairs reduction equals the
    reduction of the materialised matrix; single-pass and reference-order passes agree."""
    import groan_rs_b200 as g
    import torch
    # ---- configs[4]: 4M atoms, box 34 nm, centre + RMSD
    n, F, L = 4_000_000, 4, np.array([34.0, 34.0, 34.0], np.float32)
    scale = 5.5 / 131070.0
    masses = np.random.default_rng(3).uniform(1.0, 100.0, n).astype(np.float32)
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], np.float32)
    rot = np.stack([np.eye(3, dtype=np.float32), R, R.T, R @ R]).reshape(F, 9)
    cen = np.array([[17, 17, 17], [0.5, 33.8, 10.0], [33.9, 0.1, 0.2], [8.0, 25.0, 31.0]], np.float32)
    s = g.System(n, masses=masses, max_frames=F)
    ref = g.System(n, masses=masses)
    ref.set_frames(s.synth_blob_ref(11, scale, L / 2), L)
    idx = np.arange(n)
    s.group_create_from_indices("G", idx)
    ref.group_create_from_indices("G", idx)
    s.synth_blob(11, 0, F, scale, 0.0, rot, cen, L, wrap=True)   # noise-free rigid copies
    rotm = np.empty((F, 9), np.float32)
    c, r = s.group_center_and_rmsd(ref, "G", rot=rotm)
    # RMSD ~ 0 is below what f32 products can resolve: the cancellation guard hands every frame to the f64 passes
    assert s.fallback_frames() == F
    assert np.all(r <= 2e-4), r                                   # rigid copy: RMSD ~ 0 (f32 coordinates of a rotated blob)
    # kabsch returns r with p_rotated = r^T p (rmsd.rs:586): frames were built as x = R p + c, so r^T = R
    for f in range(F):
        assert np.abs(rotm[f].reshape(3, 3).T - rot[f].reshape(3, 3)).max() <= 1e-4, f
    # the blob is symmetric about its centre up to sampling: the PBC centre is the generator's centre (mod L)
    dc = (c - cen + L / 2) % L - L / 2
    assert np.abs(dc).max() <= 5e-3, dc
    c_sep, r_sep = s.group_get_center("G"), s.calc_rmsd(ref, "G")
    assert np.abs(c - c_sep).max() <= 8e-6 and np.abs(r - r_sep).max() <= 2e-5
    # translate every atom by integer box vectors: wrapped coordinates, centre and RMSD are unchanged
    before = torch.empty((F, n, 3), dtype=torch.float32, device="cuda")
    s.get_frames(out=before)
    s.sync()
    s.atoms_translate([34.0 * 2, -34.0, 34.0 * 3])
    after = torch.empty_like(before)
    s.get_frames(out=after)
    s.sync()
    dlt = (after - before).abs()
    dlt = torch.minimum(dlt, (dlt - 34.0).abs())                  # x = 0 may come back as x = L (closed interval, strict compares)
    assert float(dlt.max()) <= 2e-5                               # back in the same box image
    s.atoms_wrap()
    again = torch.empty_like(before)
    s.get_frames(out=again)
    s.sync()
    assert torch.equal(again, after)                              # idempotent
    c2, r2 = s.group_center_and_rmsd(ref, "G")
    assert np.abs(((c2 - c + L / 2) % L) - L / 2).max() <= 2e-5 and np.abs(r2 - r).max() <= 2e-4
    del before, after, again
    s.close()
    ref.close()
    # ---- configs[3]: 1M atoms, 2 000 x 200 000 all-pairs
    N, n1, n2 = 1_000_000, 2000, 200_000
    p = g.System(N, max_frames=1)
    p.group_create_from_indices("A", np.arange(n1))
    p.group_create_from_indices("B", np.arange(500_000, 500_000 + n2))
    p.synth_uniform(20261018, 0, 1, [-2.15] * 3, [25.8] * 3, [21.5] * 3)
    mat = torch.empty((1, n1, n2), dtype=torch.float32, device="cuda")
    p.group_all_distances("A", "B", g.Dimension.XYZ, out=mat)
    red = p.group_all_distances_reduce("A", "B", g.Dimension.XYZ, cutoff=1.0)
    p.sync()
    flat = mat.view(-1)
    assert float(flat.min()) == float(red["min"][0]) and float(flat.max()) == float(red["max"][0])
    first_min = int((flat == flat.min()).nonzero()[0])            # first minimum of the row-major scan
    last_max = int((flat == flat.max()).nonzero()[-1])            # last maximum
    assert (first_min // n2, first_min % n2) == tuple(int(v) for v in red["argmin"][0])
    assert (last_max // n2, last_max % n2) == tuple(int(v) for v in red["argmax"][0])
    assert int((flat < 1.0).sum()) == int(red["count"][0])
    assert float(flat.max()) <= 0.5 * np.sqrt(3) * 21.5 + 1e-4     # no distance beyond half the box diagonal
    # a 64 x 4096 sub-block against the ref32 oracle, bit for bit
    fr = p.get_frames()[0]
    sub = orc.all_distances(fr, np.arange(64), np.arange(500_000, 500_000 + 4096), "XYZ", [21.5] * 3)
    assert np.array_equal(bits(mat[0, :64, :4096].cpu().numpy()), bits(sub))
    # the cell-grid search finds exactly the pairs of the matrix below the cutoff (2 000 x 200 000, cutoff 1.0 and 0.3 nm):
    # the same count, every stored pair below the cutoff with the matrix's own distance, no pair twice
    for cutoff in (1.0, 0.3):
        want = int((flat < cutoff).sum())
        count, pairs, dist = p.group_pairs_within("A", "B", cutoff, capacity=want + 16, with_distances=True)
        assert int(count[0]) == want
        pi = torch.from_numpy(pairs[0, :want].astype(np.int64)).cuda()
        lin = pi[:, 0] * n2 + pi[:, 1]
        assert int(torch.unique(lin).numel()) == want
        assert torch.equal(flat[lin].cpu(), torch.from_numpy(dist[0, :want]))
        assert float(dist[0, :want].max()) < cutoff


def test_device_side_fallback_matches_exact_only():
    """A large contiguous group that is NOT compact (uniform in the box) goes through the TMA-fed single-pass kernel, which
    flags every frame and tail-launches the reference-order passes from the device; results must be bit-identical to
    GROAN_FLAG_EXACT_ONLY and to the host-launched fallback (GROAN_FLAG_HOST_FALLBACK)."""
    import groan_rs_b200 as g
    n, F = 120_000, 3
    L = np.array([11.0, 12.0, 13.0], np.float32)
    masses = np.random.default_rng(8).uniform(1.0, 50.0, n).astype(np.float32)
    res = {}
    for name, flags in (("device", 0), ("host", g.FLAG_HOST_FALLBACK), ("exact", g.FLAG_EXACT_ONLY)):
        s = g.System(n, masses=masses, max_frames=F)
        s.set_flags(flags)
        s.synth_uniform(99, 0, F, [0, 0, 0], L, L)
        ref = g.System(n, masses=masses)
        ref.set_frames(s.get_frames()[0], L)
        for x in (s, ref):
            x.group_create_from_indices("G", np.arange(16, n - 8))
        c = s.group_get_center("G")
        nf_c = s.fallback_frames()
        m = s.group_get_com("G")
        r = s.calc_rmsd(ref, "G")
        nf_r = s.fallback_frames()
        c2, r2 = s.group_center_and_rmsd(ref, "G")
        res[name] = (c, m, r, c2, r2)
        if name != "exact":
            assert nf_c == F and nf_r == F
    for name in ("device", "host"):
        for a, b in zip(res[name], res["exact"]):
            assert np.array_equal(bits(a), bits(b)), name


# ------------------------------------------------------------------ quad kernels (kernels_quad.cuh)
@pytest.mark.parametrize("n", [300_000, 300_001, 300_002])
def test_quad_kernels_match_pair_kernels_and_x64(n):
    """kernels_quad.cuh (quads of atoms on the TMA-fed ring, permuted reference, sine-only image decision) against the gather
    kernels (GROAN_FLAG_NO_TMA: an independent implementation of the same sums) and the exact64 oracle: heads of 0..3 atoms, ragged tails, a group smaller than one chunk per
    CTA, blobs straddling box faces (frame 0 of _blob_system sits on two of them), non-cubic box, and frame sizes that
    are not multiples of four atoms (the group's 16-byte phase then changes from frame to frame: one permuted reference
    per phase)."""
    import groan_rs_b200 as g
    F, L = 6, np.array([20.0, 23.0, 17.5], np.float32)
    masses = np.random.default_rng(4).uniform(1.0, 100.0, n).astype(np.float32)
    scale, nscale = 3.0 / 131070.0, 0.03 / 37837.23
    quad = _blob_system(n, F, L, 11, scale, nscale, masses)
    pair = _blob_system(n, F, L, 11, scale, nscale, masses, flags=g.FLAG_NO_TMA)
    ref = g.System(n, masses=masses)
    ref_xyz = quad.synth_blob_ref(11, scale, L / 2)
    ref.set_frames(ref_xyz, L)
    other = np.random.default_rng(5).uniform(1.0, 100.0, n).astype(np.float32)  # reference masses != target masses
    ref2 = g.System(n, masses=other)
    ref2.set_frames(ref_xyz, L)
    frames = quad.get_frames()
    groups = {"h0": np.arange(0, n), "h3": np.arange(1, n - 2), "h2": np.arange(2, n - 5), "h1": np.arange(3, 9000),
              "small": np.arange(40_000, 45_001)}
    for name, idx in groups.items():
        for s in (quad, pair, ref, ref2):
            s.group_create_from_indices(name, idx)
        cq, cp = quad.group_get_center(name), pair.group_get_center(name)
        assert quad.fallback_frames() == 0 and pair.fallback_frames() == 0
        mq, mp = quad.group_get_com(name), pair.group_get_com(name)
        rq, rp = quad.calc_rmsd(ref, name), pair.calc_rmsd(ref, name)
        assert quad.fallback_frames() == 0
        assert np.abs(cq - cp).max() <= 4e-6 and np.abs(mq - mp).max() <= 4e-6 and np.abs(rq - rp).max() <= 2e-6
        for weighted, sep in ((False, cq), (True, mq)):
            c2, r2 = quad.group_center_and_rmsd(ref, name, weighted=weighted)
            assert quad.fallback_frames() == 0
            assert np.abs(c2 - sep).max() <= 4e-6 and np.abs(r2 - rq).max() <= 2e-6
        # SAME_MASS = false instantiations (weights from the reference system, COM of the target from its own masses:
        # rmsd.rs:154,192); the x64 oracle takes one mass array, so these are pinned to the pair kernels only
        c3, r3 = quad.group_center_and_rmsd(ref2, name, weighted=True)
        c4, r4 = pair.group_center_and_rmsd(ref2, name, weighted=True)
        assert np.abs(c3 - c4).max() <= 4e-6 and np.abs(r3 - r4).max() <= 2e-6
        for f in (0, F - 1):
            c64 = orc.get_center_x64(frames[f], idx, L)
            m64 = orc.get_center_x64(frames[f], idx, L, mass=masses[idx])
            r64, _ = orc.calc_rmsd_x64(ref_xyz, idx, L, masses[idx], frames[f], idx, L)
            assert np.abs(cq[f] - c64).max() <= TOL_CENTER and np.abs(mq[f] - m64).max() <= TOL_CENTER, (name, f, cq[f], c64)
            assert abs(rq[f] - r64) <= 2e-5, (name, f, rq[f], r64)


def test_fused_centre_second_tier():
    """The fused centre + RMSD kernel decides the centre's periodic image from the moments sum d, sum d^2
    (finish_center_moments, kernels_quad.cuh).  Frames whose mean lies too close to a box face for that go through the
    sine-sum centre pass (second tier): launched from the host behind every fused launch for up to four frames, from the
    device for what exceeds that, or gated by per-frame flags under GROAN_FLAG_HOST_FALLBACK; none of these is a fallback to
    the exact passes.  All frames must match the exact64 oracle and the separate calls, and the three ways must agree bit for bit."""
    import groan_rs_b200 as g
    n, F = 262_144 + 4, 12
    L = np.array([12.0, 11.0, 13.0], np.float32)
    masses = np.random.default_rng(21).uniform(1.0, 100.0, n).astype(np.float32)
    scale, nscale = 1.3 / 131070.0, 0.02 / 37837.23
    rot = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (F, 1))
    cen = np.array([[6.0, 5.5, 6.5],       # inside the box: no face straddled, nothing to decide
                    [0.5, 5.5, 6.5],       # straddles x = 0 with the mean 0.5 nm inside: the moments certify it
                    [0.003, 5.5, 6.5],     # mean 3 pm from the face: second tier
                    [11.998, 5.5, 6.5],    # the same from the other side
                    [6.0, 10.997, 0.004],  # two faces at once
                    [11.5, 10.5, 12.4],    # three faces straddled, all certified by the moments
                    [6.0, 5.5, 12.9999],   # 0.1 pm: the sine sum cannot decide either -> exact passes (third tier)
                    [3.0, 3.0, 3.0],
                    [6.0, 0.002, 6.5],     # frames 8 .. 10: more second-tier frames than the host-launched pass has room for
                    [6.0, 5.5, 0.0025],
                    [0.0015, 5.5, 6.5],
                    [9.0, 9.0, 9.0]], np.float32)
    idx = np.arange(2, n - 1, dtype=np.uint32)
    res = {}
    for name, flags, nf in (("device", 0, F), ("host", g.FLAG_HOST_FALLBACK, F), ("few", 0, 8)):
        s = g.System(n, masses=masses, max_frames=nf)
        s.set_flags(flags)
        s.synth_blob(33, 0, nf, scale, nscale, rot[:nf], cen[:nf], L, wrap=True)
        ref = g.System(n, masses=masses)
        ref_xyz = s.synth_blob_ref(33, scale, L / 2)
        ref.set_frames(ref_xyz, L)
        for x in (s, ref):
            x.group_create_from_indices("G", idx)
        out = {}
        for weighted in (False, True):
            c2, r2 = s.group_center_and_rmsd(ref, "G", weighted=weighted)
            second, slow = s.second_pass_frames(), s.fallback_frames()
            assert second == (7 if nf == F else 4), (name, weighted, second)   # frames 2, 3, 4, 6 (, 8, 9, 10)
            assert slow <= 1, (name, weighted, slow)                             # frame 6 only, if the sine sum gives up on it
            sep = s.group_get_com("G") if weighted else s.group_get_center("G")
            assert np.abs(c2 - sep).max() <= 4e-6, (name, weighted, c2, sep)
            c3, r3 = s.group_center_and_rmsd(ref, "G", weighted=weighted)        # the counters were re-armed
            assert np.array_equal(bits(c2), bits(c3)) and np.array_equal(bits(r2), bits(r3))
            out[weighted] = (c2, r2)
        res[name] = out
        if name == "device":
            frames = s.get_frames()
            for f in range(nf):
                e64 = orc.get_center_x64(frames[f], idx, L)
                assert np.abs(out[False][0][f] - e64).max() <= 1e-5, (f, out[False][0][f], e64)
                r64, _ = orc.calc_rmsd_x64(ref_xyz, idx, L, masses[idx], frames[f], idx, L)
                assert abs(out[False][1][f] - r64) <= 1e-4
    for w in (False, True):
        assert np.array_equal(bits(res["device"][w][0]), bits(res["host"][w][0]))
        assert np.array_equal(bits(res["device"][w][1]), bits(res["host"][w][1]))


def test_back_to_back_calls_keep_stream_order():
    """Back-to-back calls that write the SAME result buffers -- one on a group every frame of which needs the device-launched
    fallback passes (tail-launched grids), one on a compact group -- must leave the results of the call issued last, and
    repeated calls must be bit-identical (fixed-order reductions, static chunk assignment)."""
    import groan_rs_b200 as g
    import torch
    n, F = 200_000, 4
    L = np.array([12.0, 12.0, 12.0], np.float32)
    masses = np.random.default_rng(3).uniform(1.0, 50.0, n).astype(np.float32)
    out = {}
    for name, flags in (("device", 0), ("host", g.FLAG_HOST_FALLBACK)):
        s = g.System(n, masses=masses, max_frames=F)
        s.set_flags(flags)
        s.synth_uniform(5, 0, F, [0, 0, 0], L, L)
        ref = g.System(n, masses=masses)
        ref.set_frames(s.get_frames()[0], L)
        for x in (s, ref):
            x.group_create_from_indices("wide", np.arange(0, n))            # uniform in the box: every frame is flagged
            x.group_create_from_indices("narrow", np.arange(1000, 1000 + 8192))
        # make the narrow group compact: its atoms are uniform in the box too, so move them into a small cube
        fr = s.get_frames().copy()
        fr[:, 1000:1000 + 8192] = 5.0 + 0.1 * fr[:, 1000:1000 + 8192]
        s.set_frames(fr, np.tile(L, (F, 1)))
        rf = fr[0].copy()  # reference = frame 0 plus a ripple, so that no RMSD is ~0 (that would be flagged, by design)
        rf[1000:1000 + 8192] += (0.02 * np.sin(np.arange(8192 * 3, dtype=np.float32))).reshape(8192, 3)
        ref.set_frames(rf, L)
        dev = torch.device("cuda", 0)
        d_c = torch.empty((F, 3), dtype=torch.float32, device=dev)
        d_r = torch.empty((F,), dtype=torch.float32, device=dev)
        want_wide = tuple(np.array(a, copy=True) for a in s.group_center_and_rmsd(ref, "wide"))
        want_narrow = tuple(np.array(a, copy=True) for a in s.group_center_and_rmsd(ref, "narrow"))
        assert s.fallback_frames() == 0
        for rep in range(60):
            s.group_center_and_rmsd(ref, "wide", center_out=d_c, rmsd_out=d_r)
            s.group_center_and_rmsd(ref, "narrow", center_out=d_c, rmsd_out=d_r)
            if rep % 20 == 19:
                s.sync()
                assert np.array_equal(bits(d_c.cpu().numpy()), bits(want_narrow[0])), (name, rep)
                assert np.array_equal(bits(d_r.cpu().numpy()), bits(want_narrow[1])), (name, rep)
        for rep in range(20):
            s.group_get_center("narrow", out=d_c)
            s.group_center_and_rmsd(ref, "wide", center_out=d_c, rmsd_out=d_r)
        s.sync()
        assert np.array_equal(bits(d_c.cpu().numpy()), bits(want_wide[0])), name
        assert np.array_equal(bits(d_r.cpu().numpy()), bits(want_wide[1])), name
        out[name] = (want_wide, want_narrow)
    for a, b in zip(out["device"][0] + out["device"][1], out["host"][0] + out["host"][1]):
        assert np.array_equal(bits(a), bits(b))


# ------------------------------------------------------------------ whole groups / molecules, centering
def test_make_whole_goldens_gpu(conect):
    """modifying.rs:1108-1153: conect.pdb translated by (3.5, 4.5, -3.0), then make_group_whole("all") /
    make_molecules_whole(): the reference's expected .gro files (3 decimals) and the ref32 oracle"""
    from conftest import mol_refs
    xyz, box, n = conect["xyz"], conect["box"], conect["xyz"].shape[0]
    s = _sys(n)
    s.set_frames(xyz, box)
    s.atoms_translate(conect["translate"])
    moved = s.get_frames()[0].copy()
    assert np.array_equal(bits(moved), bits(orc.translate(xyz, np.arange(n), conect["translate"], box)[0]))
    s.make_group_whole("all")
    got = s.get_frames()[0]
    assert np.abs(got - conect["whole_group"]).max() < 5.1e-4
    assert np.abs(got - orc.make_group_whole(moved, np.arange(n), box)).max() <= TOL_CENTER
    s.set_frames(moved, box)
    s.add_bonds(conect["bonds"])
    s.make_molecules_whole()
    got = s.get_frames()[0]
    assert np.abs(got - conect["whole_molecules"]).max() < 5.1e-4
    # per-atom arithmetic is the reference's, f32: bit-identical to the oracle
    assert np.array_equal(bits(got), bits(orc.make_molecules_whole(moved, mol_refs(n, conect["bonds"]), box)))


@pytest.mark.parametrize("dim", ["None", "X", "Y", "Z", "XY", "XZ", "YZ", "XYZ"])
def test_atoms_center_gpu(example, dim):
    """utility.rs:336-520 through the oracle (pinned to those numbers by tests/test_oracle_kat.py): centre and COM variants"""
    xyz, box = example["xyz"], example["box"]
    n = xyz.shape[0]
    m = np.random.default_rng(6).uniform(1.0, 80.0, n).astype(np.float32)
    s = _sys(n, masses=m, max_frames=2)
    s.group_create_from_indices("Protein", example["Protein"])
    for weighted in (False, True):
        s.set_frames(np.stack([xyz, xyz + np.float32(0.37)]), np.stack([box, box]))
        (s.atoms_center_mass if weighted else s.atoms_center)("Protein", _dim(dim))
        got = s.get_frames()
        for f, src in enumerate((xyz, xyz + np.float32(0.37))):
            exp = orc.atoms_center(src, example["Protein"], dim, box, mass=m[example["Protein"]] if weighted else None)
            L = box.diagonal()
            d = np.abs(got[f] - exp)
            d = np.minimum(d, np.abs(d - L))  # an atom within 1e-6 of a box face may land on either side of it
            assert d.max() <= TOL_CENTER, (dim, weighted, f, d.max())


def test_make_whole_synthetic_molecules():
    """many small molecules scattered over the box, several frames, atoms shifted by random whole box vectors: every molecule
    comes out whole (bit-identical to the oracle), free atoms are untouched, and the group version agrees with the oracle"""
    rng = np.random.default_rng(12)
    n_mol, per = 20_000, 5
    n = n_mol * per + 777  # + free atoms
    L = np.array([11.0, 12.5, 9.75], np.float32)
    F = 3
    frames = np.empty((F, n, 3), np.float32)
    for f in range(F):
        start = rng.uniform(0, L, size=(n_mol, 1, 3))
        mol = start + np.cumsum(rng.normal(0, 0.12, size=(n_mol, per, 3)), axis=1)
        free = rng.uniform(-0.2 * L, 1.2 * L, size=(777, 3))
        x = np.concatenate([mol.reshape(-1, 3), free]).astype(np.float32)
        x[: n_mol * per] += (rng.integers(-2, 3, size=(n_mol * per, 3)) * L).astype(np.float32)
        frames[f] = x
    bonds = np.stack([np.arange(n_mol * per - 1), np.arange(1, n_mol * per)], axis=1)
    bonds = bonds[(bonds[:, 1] % per) != 0]
    from conftest import mol_refs
    ref = mol_refs(n, bonds)
    s = _sys(n, max_frames=F)
    s.add_bonds(bonds)
    s.set_frames(frames, np.tile(L, (F, 1)))
    s.make_molecules_whole()
    got = s.get_frames()
    for f in range(F):
        exp = orc.make_molecules_whole(frames[f], ref, L)
        assert np.array_equal(bits(got[f]), bits(exp)), f
        m = got[f, : n_mol * per].reshape(n_mol, per, 3)
        assert np.abs(m - m[:, :1]).max() < 2.0  # whole: every atom within a few bond lengths of its reference atom
        assert np.all((m[:, 0] >= 0) & (m[:, 0] <= L))  # reference atoms inside the box
    idx = np.arange(40, 40 + per * 300)
    s.group_create_from_indices("blob", idx)
    small = frames.copy()
    small[:, idx] = (np.float32(3.0) + 0.2 * frames[:, idx] / L).astype(np.float32) + \
        (rng.integers(-1, 2, size=(F, len(idx), 3)) * L).astype(np.float32)
    s.set_frames(small, np.tile(L, (F, 1)))
    s.make_group_whole("blob")
    got = s.get_frames()
    for f in range(F):
        exp = orc.make_group_whole(small[f], idx, L)
        assert np.abs(got[f] - exp).max() <= TOL_CENTER
        assert np.ptp(got[f, idx], axis=0).max() < 1.2  # the blob spans ~1 nm; before, its atoms were up to a box apart
        other = np.setdiff1d(np.arange(n), idx)
        assert np.array_equal(bits(got[f, other]), bits(small[f, other]))


def test_whole_and_center_error_paths(protein):
    """error order of the 8f rank-1 functions: modifying.rs:1155-1250 (make_*_whole_fail_*), utility.rs:452-520"""
    import groan_rs_b200 as g
    s = _sys(61, max_frames=2)
    s.group_create_from_indices("Protein", range(61))
    s.group_create_from_indices("Empty", [])
    s.set_frames(protein["frames"][:2], None)  # no box
    for call in (lambda: s.make_group_whole("Protein"), lambda: s.make_molecules_whole(),
                 lambda: s.atoms_center("Protein", g.Dimension.XYZ)):
        with pytest.raises(g.GroanError) as ei:
            call()
        assert "DoesNotExist" in ei.value.variant
    s.set_frames(protein["frames"][:2], protein["boxes"][:2])
    with pytest.raises(g.GroupError) as ei:
        s.make_group_whole("Nope")
    assert ei.value.variant == "NotFound"
    with pytest.raises(g.GroupError) as ei:
        s.atoms_center("Empty")
    assert ei.value.variant == "EmptyGroup"
    with pytest.raises(g.GroupError) as ei:  # no masses
        s.atoms_center_mass("Protein")
    assert "NoMass" in ei.value.variant
    valid = np.ones((2, 61), np.uint8)
    valid[1, 5] = 0
    s.set_valid(valid)
    with pytest.raises(g.GroupError) as ei:
        s.make_group_whole("Protein")
    assert "NoPosition(5)" in ei.value.variant
    s.add_bonds([(4, 5), (5, 6)])
    with pytest.raises(g.GroanError) as ei:
        s.make_molecules_whole()
    assert "NoPosition(5)" in ei.value.variant
    s.group_create_from_indices("Head", range(5))
    with pytest.raises(g.GroupError) as ei:  # the reference group is fine, but atoms_translate needs every atom
        s.atoms_center("Head")
    assert "NoPosition(5)" in ei.value.variant
    s.add_bonds([(0, 1)])  # atom 5 is now a free atom: make_molecules_whole does not look at it (modifying.rs:267-270)
    s.make_molecules_whole()
    with pytest.raises(g.GpuError):  # a reference atom must be the lowest index of its molecule
        s._check(s._lib.groan_gpu_set_molecules(s._h, np.array([1, 1] + [0xFFFFFFFF] * 59, np.uint32).ctypes.data_as(
            __import__("ctypes").c_void_p)), "set_molecules")
    # triclinic boxes are rejected like everywhere else in the reference (simbox.rs:230-236)
    tb = protein["boxes"][:2].copy()
    tb[:, 1, 0] = 1.0
    s.set_valid(None)
    s.set_frames(protein["frames"][:2], tb)
    for call in (lambda: s.make_group_whole("Protein"), lambda: s.make_molecules_whole(), lambda: s.atoms_center("Protein")):
        with pytest.raises(g.GroanError) as ei:
            call()
        assert "NotOrthogonal" in ei.value.variant


# ------------------------------------------------------------------ frames uploaded as the xtc decoder's integers
def test_quantized_frames_are_the_readers_floats(example, short_traj):
    """groan_gpu_push_frames_quantized: short_trajectory.xtc as int16 lattice points (tests/golden, proved by
    oracle/gen_golden.py to reproduce read_xtc's floats bit for bit) against the same frames pushed as f32: the frames on
    the device and every result are bit-identical; int32 and a per-frame origin give the same again."""
    from conftest import load_golden
    g0 = load_golden("short_trajectory")
    q, prec = g0["q"], float(g0["prec"])
    frames, boxes = short_traj["frames"], short_traj["boxes"]
    F, n = q.shape[0], q.shape[1]
    m = np.ones(n, np.float32)
    m[:61] = short_traj["protein_mass"]
    ref = _sys(n, masses=m)
    ref.group_create_from_indices("Protein", example["Protein"])
    ref.set_frames(example["xyz"], example["box"].reshape(1, 9))
    res = []
    org = q.reshape(F, -1, 3).min(axis=1).astype(np.int32)
    variants = (("f32", None), ("i16", (q, None)), ("i32", (q.astype(np.int32), None)),
                ("i16+origin", ((q.astype(np.int32) - org[:, None, :]).astype(np.int16), org)))
    for name, v in variants:
        s = _sys(n, masses=m, max_frames=F)
        s.group_create_from_indices("Protein", example["Protein"])
        if v is None:
            s.set_frames(frames, boxes)
        else:
            s.set_frames_quantized(v[0], prec, boxes, origin=v[1])
        got = s.get_frames()
        assert np.array_equal(bits(got), bits(frames)), name
        res.append((s.group_get_center("Protein"), s.calc_rmsd(ref, "Protein")))
    for c, r in res[1:]:
        assert np.array_equal(bits(c), bits(res[0][0])) and np.array_equal(bits(r), bits(res[0][1]))
    kat = [0.23669721, 0.2634763, 0.26021627, 0.21364464, 0.22166993, 0.19383307, 0.26422343, 0.27013618, 0.26398134,
           0.23475659, 0.24208021]  # rmsd.rs:811-814
    assert np.abs(res[1][1] - np.array(kat, np.float32)).max() < TOL_RMSD
    # the encoder's direction: what xtc writing would store for the frames it has just read is what was read
    assert np.array_equal(s.get_frames_quantized(prec), q.astype(np.int32))
    # ... and for the fitted trajectory (rmsd.rs:952-994, golden short_trajectory_fit.xtc): the same lattice points as the
    # reference wrote, up to the fit's own 7e-5 nm pushing a coordinate across a rounding boundary
    s.calc_rmsd_and_fit(ref, "Protein")
    fq = s.get_frames_quantized(prec)
    diff = np.abs(fq - g0["fit_q"].astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() > 0.97, (diff.max(), (diff == 0).mean())


# ------------------------------------------------------------------ cutoff pair search through a cell grid
@pytest.mark.parametrize("cutoff", [0.35, 1.1, 4.0, 30.0])
def test_pairs_within_matches_brute_force(cutoff):
    """CellGrid + neighbors_iter + distance filter (cellgrid.rs:301-420) against the plain double loop: the same SET of pairs
    per frame, distances bit-identical to the all-pairs matrix; cutoffs from a fraction of a cell to larger than the box
    (grids of 1 and 2 cells per axis), atoms outside the box, boxes that change from frame to frame."""
    rng = np.random.default_rng(21)
    n, F = 6000, 3
    L = np.array([[7.5, 8.25, 6.75], [7.4, 8.3, 6.9], [7.6, 8.2, 6.7]], np.float32)
    frames = np.stack([rng.uniform(-0.15 * L[f], 1.15 * L[f], size=(n, 3)) for f in range(F)]).astype(np.float32)
    frames[:, 10] = frames[:, 11]            # coincident atoms: distance 0
    frames[0, 12] = [0.0, L[0, 1], 3.0]      # on the box faces
    frames[1, 213] = [40.0, -30.0, 100.0]   # many box lengths away: this frame takes the reference's loop form of min_image
    s = _sys(n, max_frames=F)
    a_idx, b_idx = np.arange(0, 300), np.arange(200, 5200)
    s.group_create_from_indices("A", a_idx)
    s.group_create_from_indices("B", b_idx)
    s.set_frames(frames, L)
    exp = [orc.pairs_within(frames[f], a_idx, b_idx, cutoff, L[f], capacity=300 * 5000) for f in range(F)]
    cap = max(e[0] for e in exp) + 5
    count, pairs, dist = s.group_pairs_within("A", "B", cutoff, capacity=cap, with_distances=True)
    for f in range(F):
        ec, ep, ed = exp[f]
        assert int(count[f]) == ec, (cutoff, f, int(count[f]), ec)
        got = {(int(i), int(j)): d for (i, j), d in zip(pairs[f, :ec], dist[f, :ec])}
        assert len(got) == ec  # no pair twice
        want = {(int(i), int(j)): d for (i, j), d in zip(ep, ed)}
        assert got.keys() == want.keys()
        assert all(np.float32(got[k]).view(np.uint32) == np.float32(want[k]).view(np.uint32) for k in want)
    # count only, and a capacity smaller than the number of pairs: the count stays complete
    c2, _, _ = s.group_pairs_within("A", "B", cutoff)
    assert np.array_equal(c2, count)
    c3, p3, _ = s.group_pairs_within("A", "B", cutoff, capacity=7)
    assert np.array_equal(c3, count)
    for f in range(F):
        want = {(int(i), int(j)) for i, j in exp[f][1]}
        assert all((int(i), int(j)) in want for i, j in p3[f, :min(7, exp[f][0])])


def test_pairs_within_self_search_and_errors(protein):
    """a group against itself (both (i, j) and (j, i), and the diagonal, like the all-pairs matrix) == count of the fused
    all-pairs reduction; error order of CellGrid::new (cellgrid.rs:308-330)"""
    import groan_rs_b200 as g
    rng = np.random.default_rng(5)
    n = 20_000
    L = np.array([9.0, 9.0, 9.0], np.float32)
    x = rng.uniform(0, L, size=(n, 3)).astype(np.float32)
    s = _sys(n)
    s.group_create_from_indices("G", np.arange(n))
    s.set_frames(x, L)
    count, _, _ = s.group_pairs_within("G", "G", 0.6)
    red = s.group_all_distances_reduce("G", "G", g.Dimension.XYZ, cutoff=0.6)
    assert int(count[0]) == int(red["count"][0]) and int(count[0]) > n
    e = _sys(61, max_frames=1)
    e.group_create_from_indices("P", range(61))
    e.group_create_from_indices("Empty", [])
    e.set_frames(protein["frames"][:1], None)
    with pytest.raises(g.GroanError) as ei:
        e.group_pairs_within("P", "P", 1.0)
    assert "DoesNotExist" in ei.value.variant
    e.set_frames(protein["frames"][:1], protein["boxes"][:1])
    with pytest.raises(g.GroupError):
        e.group_pairs_within("P", "Nope", 1.0)
    c, _, _ = e.group_pairs_within("P", "Empty", 1.0)
    assert int(c[0]) == 0
    with pytest.raises(g.GpuError):
        e.group_pairs_within("P", "P", -1.0)


# ------------------------------------------------------------------ round 2: the benchmarked geometry, odd sizes, Kabsch KATs
def _bench_case(n, F, frame0=0):
    """the bench.py workload (BASELINE configs[4]): same generator, seed, masses and per-frame rigid motions"""
    import bench
    import groan_rs_b200 as g
    m = np.random.default_rng(bench.SEED + 1).uniform(1.0, 100.0, n).astype(np.float32)
    L = np.array([bench.BOX] * 3, np.float32)
    s = g.System(n, masses=m, max_frames=F)
    ref = g.System(n, masses=m, max_frames=1)
    idx = np.arange(n, dtype=np.uint32)
    s.group_create_from_indices("G", idx)
    ref.group_create_from_indices("G", idx)
    ref_xyz = s.synth_blob_ref(bench.SEED, bench.BLOB_SCALE, [bench.BOX / 2] * 3)
    ref.set_frames(ref_xyz, L)
    rot, cen = bench.frame_params(frame0, F)
    s.synth_blob(bench.SEED, frame0, F, bench.BLOB_SCALE, bench.NOISE_SCALE, rot, cen, L, wrap=True)
    return s, ref, ref_xyz, m, idx, L, rot, cen


def _check_sampled_frames(s, ref, ref_xyz, m, idx, L, rot, cen, frame0, sample, expect_no_fallback=True):
    """SURVEY 8d cfg5 protocol: fused call, group_get_center and calc_rmsd against the exact64 oracle on sampled frames
    (centre <= 1e-5 nm, RMSD <= 1e-4 nm); prints the drift of the reference's own sequential f32 sums (ref32)."""
    import bench
    n = len(idx)
    c_f, r_f = s.group_center_and_rmsd(ref, "G")
    nf_f = s.fallback_frames()
    c_s = s.group_get_center("G")
    nf_c = s.fallback_frames()
    r_s = s.calc_rmsd(ref, "G")
    nf_r = s.fallback_frames()
    if expect_no_fallback:
        assert (nf_f, nf_c, nf_r) == (0, 0, 0), (nf_f, nf_c, nf_r)  # the single-pass kernels certified every frame: the timed path
    drift_c, drift_r = 0.0, 0.0
    for f in sample:
        fr = orc.synth_blob_frame(n, bench.SEED, frame0 + f, bench.BLOB_SCALE, bench.NOISE_SCALE, rot[f], cen[f], L, wrap=True)
        c64 = orc.get_center_x64(fr, idx, L)
        r64, _ = orc.calc_rmsd_x64(ref_xyz, idx, L, m, fr, idx, L)
        for name, c, r in (("fused", c_f, r_f), ("separate", c_s, r_s)):
            assert np.abs(c[f] - c64).max() <= TOL_CENTER, (name, f, c[f], c64)
            assert abs(float(r[f]) - float(r64)) <= TOL_RMSD, (name, f, r[f], r64)
        assert abs(float(r64) - 0.0866) < 1e-3
        c32 = orc.get_center(fr, idx, L)
        r32, _ = orc.calc_rmsd(ref_xyz, idx, L, m, fr, idx, L)
        drift_c = max(drift_c, float(np.abs(c32 - c64).max()))
        drift_r = max(drift_r, abs(float(r32) - float(r64)))
    print("n=%d F=%d: ref32 (sequential f32) drift against exact64 on %d frames: centre %.2e nm, RMSD %.2e nm"
          % (n, len(c_f), len(sample), drift_c, drift_r))


def test_benchmarked_geometry_against_exact64():
    """bench.py's timed configuration exactly -- N = G = 4 000 000, F = 37, noisy blob, group = all: 8 CTAs per frame, each
    thread folds ~1 950 atoms into its f32 partials -- against the exact64 oracle on 8 sampled frames."""
    s, ref, ref_xyz, m, idx, L, rot, cen = _bench_case(4_000_000, 37)
    _check_sampled_frames(s, ref, ref_xyz, m, idx, L, rot, cen, 0, [0, 5, 11, 17, 22, 28, 33, 36])
    s.close()
    ref.close()


@pytest.mark.parametrize("n", [4_000_001, 4_000_002, 4_000_003])
def test_odd_system_sizes_against_exact64(n):
    """frame sizes that are not a multiple of 4 atoms: the 16-byte phase of the group differs from frame to frame"""
    s, ref, ref_xyz, m, idx, L, rot, cen = _bench_case(n, 8)
    _check_sampled_frames(s, ref, ref_xyz, m, idx, L, rot, cen, 0, [0, 1, 2, 3, 7])
    s.close()
    ref.close()


def test_kabsch_kats_through_the_gpu():
    """rmsd.rs:618-780 (synthetic Kabsch cases) through groan_gpu_rmsd: the points are placed in a 50 nm box (the RMSD and the
    rotation do not depend on the translation); expected rotations are nalgebra column arrays, hence the transposes."""
    e = np.eye(3, dtype=np.float32)
    cols = lambda mm: np.array(mm, np.float32).T  # noqa: E731
    rz = cols([[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    cases = [
        (e, e, np.eye(3), 0.0),
        (e, [[0.6666667, 1.0, 0.0], [-0.3333333, 0.0, 0.0], [0.6666667, 0.0, 1.0]], rz, 0.0),
        (e, [[2, 1, 1], [1, 2, 1], [1, 1, 2]], np.eye(3), 0.0),
        (e, [[1.6666666, 2.0, 1.0], [0.6666666, 1.0, 1.0], [1.6666666, 1.0, 2.0]], rz, 0.0),
        ([[4.3, 2.1, -5.2], [1.4, 2.1, 3.9], [2.4, -3.3, 1.8]], [[2.2, 0.0, 4.6], [-1.4, 0.2, 0.3], [1.3, 9.9, 11.3]],
         cols([[0.8842437, -0.10340805, -0.45543456], [0.2840647, -0.65496445, 0.70023507], [-0.37070346, -0.7485511, -0.5497733]]),
         4.471225),
    ]
    off, L = np.float32(20.0), np.array([50.0, 50.0, 50.0], np.float32)
    for k, (p, q, r_exp, rm_exp) in enumerate(cases):
        p, q = np.asarray(p, np.float32) + off, np.asarray(q, np.float32) + off
        ref = _sys(3, masses=np.ones(3, np.float32))
        s = _sys(3, masses=np.ones(3, np.float32))
        for x in (ref, s):
            x.group_create_from_indices("g", [0, 1, 2])
        ref.set_frames(p, L)
        s.set_frames(q, L)
        rot = np.empty((1, 9), np.float32)
        rm = s.calc_rmsd(ref, "g", rot=rot)
        e_rm, e_r = orc.calc_rmsd(p, [0, 1, 2], L, np.ones(3, np.float32), q, [0, 1, 2], L)
        assert abs(float(rm[0]) - rm_exp) <= TOL_RMSD and abs(float(rm[0]) - float(e_rm)) <= TOL_RMSD, (k, rm, rm_exp, e_rm)
        # three points: H has rank <= 2 and the reflection fix decides the third axis (rmsd.rs:577-585)
        assert np.abs(rot[0].reshape(3, 3) - r_exp).max() <= 2e-4, (k, rot[0].reshape(3, 3), r_exp)


def test_reference_cache_follows_group_and_frame_changes(protein):
    """re-creating the target group drops the device-side reference; changing the reference System in place,
    or re-creating ITS group, must re-upload it"""
    m = protein["mass"]
    box = protein["box"].reshape(1, 9)
    ref = _sys(61, masses=m)
    ref.group_create_from_indices("P", range(61))
    ref.set_frames(protein["xyz"], box)
    s = _sys(61, masses=m, max_frames=11)
    s.group_create_from_indices("P", range(61))
    s.set_frames(protein["frames"], protein["boxes"])
    r0 = s.calc_rmsd(ref, "P")
    s.group_create_from_indices("P", range(61))        # same atoms: set_group drops the reference on the device
    assert np.array_equal(bits(s.calc_rmsd(ref, "P")), bits(r0))
    for x in (s, ref):                                  # another group under the same name, on both systems
        x.group_create_from_indices("P", range(10, 50))
    r1 = s.calc_rmsd(ref, "P")
    idx = np.arange(10, 50)
    for f in (0, 10):
        e, _ = orc.calc_rmsd(protein["xyz"], idx, protein["box"], m[idx], protein["frames"][f], idx, protein["boxes"][f])
        assert abs(float(r1[f]) - float(e)) <= TOL_RMSD
    ref.group_translate("P", [0.4, -0.2, 0.1])          # in-place change of the reference System: RMSD is invariant ...
    assert np.abs(s.calc_rmsd(ref, "P") - r1).max() <= 2e-5
    moved = protein["xyz"].copy()
    moved[10:50:2] += np.float32(0.3)                    # ... but a deformed reference is not
    ref.set_frames(moved, box)
    r2 = s.calc_rmsd(ref, "P")
    e, _ = orc.calc_rmsd(moved, idx, protein["box"], m[idx], protein["frames"][3], idx, protein["boxes"][3])
    assert abs(float(r2[3]) - float(e)) <= TOL_RMSD and abs(float(r2[3]) - float(r1[3])) > 1e-3


def test_builtin_all_group_has_masses_and_rmsd(protein):
    """'all' / 'All' exist from the start with the atoms' masses (System::new): com, centering and RMSD work on them"""
    m = protein["mass"]
    L = protein["box"].diagonal()
    ref = _sys(61, masses=m)
    ref.set_frames(protein["xyz"], protein["box"].reshape(1, 9))
    s = _sys(61, masses=m, max_frames=11)
    s.set_frames(protein["frames"], protein["boxes"])
    idx = np.arange(61)
    com, est = s.group_get_com("all"), s.group_estimate_com("All")
    r = s.calc_rmsd(ref, "all")
    c2, r2 = s.group_center_and_rmsd(ref, "all", weighted=True)
    assert np.allclose(r, GOLDEN_RMSD, atol=TOL_RMSD) and np.abs(r2 - r).max() <= 2e-6 and np.abs(c2 - com).max() <= 4e-6
    for f in (0, 6):
        Lf = protein["boxes"][f].reshape(3, 3).diagonal()
        assert np.abs(com[f] - orc.get_com(protein["frames"][f], idx, m, Lf)).max() <= TOL_CENTER
        assert np.abs(est[f] - orc.estimate_center(protein["frames"][f], idx, Lf, mass=m)).max() <= 1e-4
    s.atoms_center_mass("all")
    # without masses the built-in group fails like the reference: MassError::NoMass of the first atom
    import groan_rs_b200 as g
    bare = _sys(61)
    bare.set_frames(protein["xyz"], protein["box"].reshape(1, 9))
    with pytest.raises(g.GroupError) as ei:
        bare.group_get_com("all")
    assert "NoMass(0)" in ei.value.variant
    assert np.abs(bare.group_get_center("all")[0] - orc.get_center(protein["xyz"], idx, L)).max() <= TOL_CENTER


def test_caller_provided_shift_buffers(example):
    """group_wrap / group_translate accept a caller's int8 buffer (numpy or device tensor) for the shifts"""
    import torch
    xyz, box = example["xyz"], example["box"]
    n = xyz.shape[0]
    L = box.diagonal()
    s = _sys(n)
    s.set_frames(xyz, box.reshape(1, 9))
    mine = np.full((1, n, 3), 99, np.int8)
    got = s.atoms_translate([30.0, -17.0, 5.0], shifts=mine)
    assert got is mine
    _, esh = orc.translate(xyz, np.arange(n), [30.0, -17.0, 5.0], L)
    assert np.array_equal(mine[0], esh)
    s.set_frames(xyz, box.reshape(1, 9))
    dev = torch.full((1, n, 3), 99, dtype=torch.int8, device="cuda")
    s.atoms_translate([30.0, -17.0, 5.0], shifts=dev)
    s.sync()
    assert np.array_equal(dev.cpu().numpy()[0], esh)


# ------------------------------------------------------------------ users of the cell grid: guess_bonds, HBondAnalysis
def _water_box(n_mol, L, seed):
    """rigid three-site waters at random positions and orientations in a cubic box: O, H, H per molecule"""
    rng = np.random.default_rng(seed)
    o = rng.uniform(0, L, size=(n_mol, 3))
    q = rng.normal(size=(n_mol, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                  2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(n_mol, 3, 3)
    h1 = np.array([0.0957, 0.0, 0.0]), np.array([-0.024, 0.0927, 0.0])
    xyz = np.empty((n_mol, 3, 3))
    xyz[:, 0] = o
    xyz[:, 1] = o + R @ h1[0]
    xyz[:, 2] = o + R @ h1[1]
    return np.mod(xyz.reshape(-1, 3), L).astype(np.float32)


def test_guess_bonds_matches_restatement():
    """groan_gpu_guess_bonds against the restated identify_bonds (guess.rs:427-470): the same set of (i, j) pairs in every
    frame -- water boxes dense enough for chance contacts, atoms without a radius, bonds through the box faces -- and the
    topology it assigns (mol_ref of make_molecules_whole)."""
    import groan_rs_b200 as g
    n_mol, L, F = 3000, 4.2, 3
    frames = np.stack([_water_box(n_mol, L, 100 + f) for f in range(F)])
    n = frames.shape[1]
    vdw = np.tile(np.array([0.152, 0.12, 0.12], np.float32), n_mol)
    vdw[5::97] = -1.0  # some hydrogens without a radius
    s = g.System(n, max_frames=F)
    s.set_frames(frames, [L] * 3)
    bonds, no_vdw = s.guess_bonds(vdw, 0.55)
    assert no_vdw == [int(i) + 1 for i in np.nonzero(vdw < 0)[0]]
    for f in range(F):
        exp = orc.guess_bonds(frames[f], vdw, [L] * 3, 0.55)
        assert np.array_equal(bonds[f], exp), (f, len(bonds[f]), len(exp))
        assert len(exp) > 2 * n_mol - 100  # the O-H bonds, plus chance contacts
    # the bonds of frame 0 became the topology: every water is one molecule with its oxygen as reference atom
    mol = s._mol_ref.reshape(n_mol, 3)
    intact = (vdw.reshape(n_mol, 3) >= 0).all(axis=1)
    assert (mol[intact, 1] <= np.arange(n_mol)[intact] * 3).all() and (mol[intact, 1] == mol[intact, 2]).all()
    # a larger radius factor, capacity exceeded on the first try (the call is repeated with the count it reported)
    b2, _ = s.guess_bonds(vdw, 0.8, capacity=16, assign=False)
    assert np.array_equal(b2[1], orc.guess_bonds(frames[1], vdw, [L] * 3, 0.8))
    # nobody has a radius: no bonds
    b3, nv = s.guess_bonds([None] * n, assign=False)
    assert all(len(b) == 0 for b in b3) and len(nv) == n


def test_hbonds_water_goldens_gpu():
    """hbonds.rs:505-587 through groan_gpu_hbonds: the reference's own numbers -- bonds per frame (exact) and the first / last
    bond of each of the 21 frames (1e-3) -- on aa_membrane_peptide.xtc read by the product's xtc reader; two frames in full
    against the restated analyze_single."""
    import os
    import groan_rs_b200 as g
    import hbond_goldens as hg
    here = os.path.dirname(os.path.abspath(__file__))
    xf = g.xtc.XtcFile.open(os.path.join(here, "golden", "xtc", "aa_membrane_peptide.xtc"))
    water = np.load(os.path.join(here, "golden", "aa_membrane_water.npz"))
    ow = water["OW"].astype(np.int64)
    n = int(water["n_atoms"][0])
    assert xf.n_atoms == n and xf.n_frames == 21
    s = g.System(n, max_frames=21)
    s.set_frames_xtc(xf)
    s.group_create_from_indices("OW", ow)
    s.group_create_from_indices("HW", np.sort(np.concatenate([ow + 1, ow + 2])))
    s.add_bonds(np.concatenate([np.stack([ow, ow + 1], 1), np.stack([ow, ow + 2], 1)]))
    from groan_rs_b200.hbonds import HBondAnalysis, HBondChain
    an = HBondAnalysis(s, [HBondChain("OW", "OW", "HW")], [(0, 0)], hg.MAX_DISTANCE, hg.MIN_ANGLE)
    maps = an.analyze()
    frames = s.get_frames()
    for f in range(21):
        rec = maps[f][(0, 0)]
        hg.check_frame(f, zip(rec["donor"], rec["hydrogen"], rec["acceptor"], rec["distance"], rec["angle"]))
    donors = [(int(o), [int(o) + 1, int(o) + 2]) for o in ow]
    for f in (3, 17):
        L = s._boxes[f].reshape(3, 3).diagonal().copy()
        exp = orc.hbonds_single(frames[f], ow, donors, L, hg.MAX_DISTANCE, hg.MIN_ANGLE)
        rec = maps[f][(0, 0)]
        assert [(int(a), int(b), int(c)) for a, b, c in zip(rec["donor"], rec["hydrogen"], rec["acceptor"])] == [r[:3] for r in exp]
        assert np.abs(rec["distance"] - np.array([r[3] for r in exp], np.float32)).max() <= TOL_DIST
        assert np.abs(rec["angle"] - np.array([r[4] for r in exp], np.float32)).max() <= 2e-2  # acosf near 180 degrees


def test_hbonds_two_chains_and_errors():
    """analyze_pair (hbonds.rs:214-238: acceptors of chain 1 with donors of chain 2 and the other way round), an acceptor that
    is also a donor (never its own partner), the inclusive distance and angle limits, and the constructor's checks."""
    import groan_rs_b200 as g
    from groan_rs_b200.hbonds import HBondAnalysis, HBondChain, HBondError
    L, n_mol = 3.1, 900
    xyz = _water_box(n_mol, L, 7)
    n = xyz.shape[0]
    s = g.System(n)
    s.set_frames(xyz, [L] * 3)
    o = np.arange(0, n, 3)
    a_o, b_o = o[: n_mol // 2], o[n_mol // 2:]
    for name, oo in (("A_O", a_o), ("B_O", b_o)):
        s.group_create_from_indices(name, oo)
    s.group_create_from_indices("H", np.sort(np.concatenate([o + 1, o + 2])))
    s.add_bonds(np.concatenate([np.stack([o, o + 1], 1), np.stack([o, o + 2], 1)]))
    an = HBondAnalysis(s, [HBondChain("A_O", "A_O", "H"), HBondChain("B_O", "B_O", "H")], [(0, 1), (1, 1)], 0.33, 140.0)
    m = an.analyze()[0]
    da = [(int(i), [int(i) + 1, int(i) + 2]) for i in a_o]
    db = [(int(i), [int(i) + 1, int(i) + 2]) for i in b_o]
    exp01 = orc.hbonds_single(xyz, a_o, db, [L] * 3, 0.33, 140.0) + orc.hbonds_single(xyz, b_o, da, [L] * 3, 0.33, 140.0)
    exp11 = orc.hbonds_single(xyz, b_o, db, [L] * 3, 0.33, 140.0)
    for key, exp in (((0, 1), exp01), ((1, 1), exp11)):
        rec = m[key]
        assert len(exp) > 20
        assert [(int(a), int(b), int(c)) for a, b, c in zip(rec["donor"], rec["hydrogen"], rec["acceptor"])] == [r[:3] for r in exp], key
        assert np.abs(rec["angle"] - np.array([r[4] for r in exp], np.float32)).max() <= 2e-2
    assert all(int(d) != int(a) for d, a in zip(m[(1, 1)]["donor"], m[(1, 1)]["acceptor"]))
    with pytest.raises(HBondError):
        HBondAnalysis(s, [HBondChain("A_O", "A_O", "H")], [(0, 1)], 0.3, 150.0)          # pair names a chain that does not exist
    with pytest.raises(HBondError):
        HBondAnalysis(s, [HBondChain("A_O", "A_O", "H")], [(0, 0), (0, 0)], 0.3, 150.0)  # the same pair twice
    with pytest.raises(g.GroanError) as ei:  # triclinic box: CellGrid::new rejects it (cellgrid.rs:308-312)
        s.set_frames(xyz, np.array([L, 0, 0, 0.5, L, 0, 0, 0, L], np.float32).reshape(1, 9))
        an.analyze()
    assert "NotOrthogonal" in ei.value.variant


def test_box_spanning_group_skips_the_single_pass_after_the_first_call():
    """A group that spans the box (a membrane slab) fails the single pass on every frame.  The first call finds that out (single
    pass + exact passes launched from the device); from the second call on the host goes straight to the exact passes
    (FallbackPlan::feedback).  Results are the same bits either way and equal GROAN_FLAG_EXACT_ONLY; a compact group on the same
    System is not affected."""
    import groan_rs_b200 as g
    n, F = 150_000, 5
    L = np.array([14.0, 13.0, 12.0], np.float32)
    masses = np.random.default_rng(5).uniform(1.0, 60.0, n).astype(np.float32)
    out = {}
    for name, flags in (("default", 0), ("exact", g.FLAG_EXACT_ONLY)):
        s = g.System(n, masses=masses, max_frames=F)
        s.set_flags(flags)
        s.synth_uniform(21, 0, F, [0.0, 0.0, 4.0], [L[0], L[1], 3.0], L)  # a slab: uniform in x and y, 3 nm thick in z
        ref = g.System(n, masses=masses)
        ref.set_frames(s.get_frames()[0], L)
        for x in (s, ref):
            x.group_create_from_indices("slab", np.arange(8, n - 4))
        calls = []
        for k in range(4):
            c = s.group_get_center("slab")
            nf = s.fallback_frames()
            r = s.calc_rmsd(ref, "slab")
            calls.append((c, r))
            if name == "default":
                assert nf == F, (k, nf)
        launches_before = s.launch_count()
        s.group_get_center("slab")
        per_call = s.launch_count() - launches_before
        if name == "default":
            assert per_call == 2, per_call  # k_trig_quad + k_center_quad(ext_pilot): no single pass, nothing launched from the device
        for c, r in calls[1:]:
            assert np.array_equal(bits(c), bits(calls[0][0])) and np.array_equal(bits(r), bits(calls[0][1]))
        out[name] = calls[0]
        frames = s.get_frames()
        idx = np.arange(8, n - 4)
        for f in (0, F - 1):  # z is compact, x / y span the box: compare the well-defined axis with the exact64 oracle
            assert abs(calls[0][0][f][2] - orc.get_center_x64(frames[f], idx, L)[2]) <= TOL_CENTER
    assert np.array_equal(bits(out["default"][0]), bits(out["exact"][0])) and np.array_equal(bits(out["default"][1]), bits(out["exact"][1]))


@pytest.mark.gpu
@pytest.mark.parametrize("chunks_per_cta", [1, 2, 3, 4, 5, 6, 7, 9])
def test_warp_ring_every_depth(chunks_per_cta):
    """stream_quads_warp (the per-warp cp.async ring of the RMSD kernels, kernels_quad.cuh) with 1 .. 9 chunks per CTA: fewer
    chunks than stages, exactly as many, one more (the first trip of the hot loop), ragged last chunks that end inside a warp's
    slice, inside a 128-atom reference block and on a chunk boundary, heads of 0 .. 3 atoms, and a frame size that is not a
    multiple of four atoms (one launch per 16-byte phase).  Centre, RMSD and the fused call against the exact64 oracle."""
    import groan_rs_b200 as g
    F, L = 37, np.array([14.0, 15.0, 13.0], np.float32)  # 37 frames: 8 CTAs per frame, like the bench
    body = 8 * 1024 * chunks_per_cta
    n = body + 1032 + (3 if chunks_per_cta == 9 else 0)  # the last case: a phase per frame, CTAs per frame from the phase's frame count
    masses = np.random.default_rng(9).uniform(1.0, 100.0, n).astype(np.float32)
    scale, nscale = 1.5 / 131070.0, 0.02 / 37837.23
    s = _blob_system(n, F, L, 31 + chunks_per_cta, scale, nscale, masses)
    ref = g.System(n, masses=masses)
    ref_xyz = s.synth_blob_ref(31 + chunks_per_cta, scale, L / 2)
    ref.set_frames(ref_xyz, L)
    frames = s.get_frames()
    groups = {"full": np.arange(0, body), "slice": np.arange(1, body - 1024 + 36 + 1), "block": np.arange(2, body - 300),
              "plus": np.arange(3, body + 1024 + 4)}
    for name, idx in groups.items():
        for sysm in (s, ref):
            sysm.group_create_from_indices(name, idx)
        c = s.group_get_center(name)
        r = s.calc_rmsd(ref, name)
        assert s.fallback_frames() == 0
        c2, r2 = s.group_center_and_rmsd(ref, name)
        assert np.abs(c2 - c).max() <= 4e-6 and np.abs(r2 - r).max() <= 2e-6, name
        for f in (0, 1, 2, 3, F - 1):
            c64 = orc.get_center_x64(frames[f], idx, L)
            r64, _ = orc.calc_rmsd_x64(ref_xyz, idx, L, masses[idx], frames[f], idx, L)
            assert np.abs(c2[f] - c64).max() <= TOL_CENTER, (name, f, c2[f], c64)
            assert abs(r[f] - r64) <= 2e-5 and abs(r2[f] - r64) <= 2e-5, (name, f, r[f], r2[f], r64)
    s.close()
    ref.close()
