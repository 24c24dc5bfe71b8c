"""Pin the CPU oracle (oracle/groan_oracle.c, ref32 flavour) to the reference's own golden vectors.

Every expected number below is copied from an assertion in the reference's inline tests
(file:line cited per test); inputs are the reference's fixtures decoded by oracle/gen_golden.py.
"""
import numpy as np
import pytest

from oracle import oracle as orc

B4 = [4.0, 4.0, 4.0]
B10 = [10.0, 10.0, 10.0]


def approx(a, b, eps=1e-6):
    assert abs(float(a) - float(b)) <= eps, (a, b)


# ---- vector3d.rs:1017-1037
def test_wrap_kat():
    L = np.float32(2.0)
    got = [[orc.wrap1(v, L) for v in p] for p in ([-1.0, 1.5, 3.0], [2.0, 2.2, -0.3], [-54.2, 77.8, 124.5])]
    exp = [[1.0, 1.5, 1.0], [2.0, 0.2, 1.7], [1.8, 1.8, 0.5]]
    for g, e in zip(got, exp):
        for a, b in zip(g, e):
            approx(a, b, 1e-5)
    assert orc.wrap1(2.0, 2.0) == 2.0  # strict '>' : x == L stays L
    assert orc.wrap1(0.0, 2.0) == 0.0


# ---- vector3d.rs:1040-1207
@pytest.mark.parametrize("dim,exp", [("X", 1.5), ("Y", -0.2), ("Z", -1.8), ("XY", 1.51327), ("XZ", 2.34307),
                                     ("YZ", 1.81108), ("XYZ", 2.351595), ("None", 0.0)])
def test_distance_kat(dim, exp):
    p1, p2 = [1.0, 3.9, 2.6], [3.5, 0.1, 0.4]
    approx(orc.distance(p1, p2, dim, B4), exp, 1e-5)
    sign = -1.0 if dim in ("X", "Y", "Z") else 1.0
    approx(orc.distance(p2, p1, dim, B4), sign * exp, 1e-5)


def test_distance_outofbox_kat():
    p1, p2 = [-1.0, 4.5, 2.3], [3.5, -0.5, 4.2]
    approx(orc.distance(p1, p2, "X", B4), -0.5)
    approx(orc.distance(p1, p2, "Y", B4), 1.0)
    approx(orc.distance(p1, p2, "Z", B4), -1.9)


# ---- vector3d.rs:1362-1437
@pytest.mark.parametrize("p1,p2,exp", [
    ([4, 4, 5], [5, 5, 3], [1, 1, -2]), ([3, 0, 7], [1, 2, 1], [-2, 2, 4]), ([1, 2, 5], [9, 8, 6], [-2, -4, 1]),
    ([8, 9, 2], [1, 3, 9], [3, 4, -3]), ([0, 3, 10], [10, 3, 0], [0, 0, 0])])
def test_vector_to_kat(p1, p2, exp):
    v = orc.vector_to(p1, p2, B10)
    for a, b in zip(v, exp):
        approx(a, b)


def test_vector_to_equidistant():
    v = orc.vector_to([7, 4, 3], [2, 5, 2], B10)
    approx(abs(v[0]), 5.0)
    approx(v[1], 1.0)
    approx(v[2], -1.0)


# ---- analysis.rs:488-629 (artificial systems, box 10^3)
def test_center_artificial_kats():
    one = np.array([[4.5, 3.2, 1.7]], np.float32)
    c = orc.estimate_center(one, [0], B10)
    for a, b in zip(c, [4.5, 3.2, 1.7]):
        approx(a, b, 2e-6)
    two = np.array([[4.5, 3.2, 1.7], [4.0, 2.8, 3.0]], np.float32)
    for a, b in zip(orc.estimate_center(two, [0, 1], B10), [4.25, 3.0, 2.35]):
        approx(a, b, 2e-6)
    pbc = np.array([[4.5, 3.2, 1.7], [9.8, 9.5, 3.0]], np.float32)
    for a, b in zip(orc.estimate_center(pbc, [0, 1], B10), [2.15, 1.35, 2.35]):
        approx(a, b, 2e-6)
    for a, b in zip(orc.get_center(pbc, [0, 1], B10), [2.15, 1.35, 2.35]):
        approx(a, b, 2e-6)
    five = np.array([[3.3, 0.3, 2.5], [4.3, 1.2, 9.8], [3.2, 5.6, 0.5], [0.2, 9.0, 6.6], [8.7, 5.0, 2.4]], np.float32)
    out = np.array([[3.3, 10.3, 2.5], [4.3, 1.2, -0.2], [13.2, 15.6, 0.5], [10.2, -1.0, 6.6], [-1.3, 5.0, 2.4]], np.float32)
    for pts in (five, out):
        for a, b in zip(orc.estimate_center(pts, range(5), B10), [2.634386, 9.775156, 1.1748]):
            approx(a, b, 1e-4)
    # SURVEY 8c "verified restatement" values (tighter than the reference's own epsilon)
    for a, b in zip(orc.estimate_center(five, range(5), B10), [2.6343856, 9.775155, 1.1748]):
        approx(a, b, 2e-6)


# ---- analysis.rs:846-988 (COM, artificial)
def test_com_artificial_kats():
    two = np.array([[4.5, 3.2, 1.7], [4.0, 2.8, 3.0]], np.float32)
    m2 = [12.8, 0.4]
    for a, b in zip(orc.estimate_center(two, [0, 1], B10, mass=m2), [4.485, 3.188, 1.73549]):
        approx(a, b, 1e-4)
    pbc = np.array([[4.5, 3.2, 1.7], [9.8, 9.5, 3.0]], np.float32)
    for a, b in zip(orc.estimate_center(pbc, [0, 1], B10, mass=m2), [4.4904, 3.1630, 1.7355]):
        approx(a, b, 1e-4)
    for a, b in zip(orc.get_com(pbc, [0, 1], m2, B10), [4.3575745, 3.0878792, 1.7393947]):
        approx(a, b, 2e-6)
    five = np.array([[3.3, 0.3, 2.5], [4.3, 1.2, 9.8], [3.2, 5.6, 0.5], [0.2, 9.0, 6.6], [8.7, 5.0, 2.4]], np.float32)
    m5 = [10.3, 5.4, 3.8, 10.1, 7.6]
    for a, b in zip(orc.estimate_center(five, range(5), B10, mass=m5), [1.9526, 9.7567, 1.8812]):
        approx(a, b, 1e-4)


# ---- analysis.rs:631-647, 749-763 (example.gro + index.ndx)
def test_center_real_system(example):
    xyz, box = example["xyz"], example["box"]
    prot, mem = example["Protein"], example["Membrane"]
    assert len(prot) == 61 and len(mem) == 6144
    nm, npr = orc.get_center_naive(xyz, mem), orc.get_center_naive(xyz, prot)
    for a, b in zip(nm, [6.47077, 6.52237, 5.77978]):
        approx(a, b, 1e-4)
    for a, b in zip(npr, [9.85718, 2.46213, 5.45931]):
        approx(a, b, 1e-4)
    cm, cp = orc.get_center(xyz, mem, box), orc.get_center(xyz, prot, box)
    approx(cm[2], nm[2], 1e-4)
    for a, b in zip(cp, npr):
        approx(a, b, 1e-4)
    # same-mass COM == centre (analysis.rs:992-1012)
    for a, b in zip(orc.get_com(xyz, prot, np.full(61, 1.0, np.float32), box), cp):
        approx(a, b, 1e-4)


# ---- analysis.rs:1269-1354
@pytest.mark.parametrize("dim,exp", [("X", 6.3029766), ("Y", -5.566175), ("Z", -0.32046986), ("XY", 8.408913),
                                     ("XZ", 6.311118), ("YZ", 5.5753927), ("XYZ", 8.415017), ("None", 0.0)])
def test_group_distance_kat(example, dim, exp):
    d = orc.group_distance(example["xyz"], example["Protein"], example["Membrane"], dim, example["box"])
    approx(d, exp, 1e-4)


# ---- analysis.rs:1420-1530
def test_all_distances_kats(example):
    xyz, box, prot, mem = example["xyz"], example["box"], example["Protein"], example["Membrane"]
    d = orc.all_distances(xyz, prot, prot, "XYZ", box)
    assert np.array_equal(d, d.T) and np.all(np.diag(d) == 0)
    approx(d.max(), 4.597961)
    approx(d[0, 1], 0.31040135)
    approx(d[60, 0], 4.266728)
    approx(d[60, 59], 0.31425142)
    z = orc.all_distances(xyz, prot, prot, "Z", box)
    assert np.array_equal(z, -z.T)
    approx(z.max(), 4.383, 1e-5)
    approx(z.min(), -4.383, 1e-5)
    approx(z[0, 1], 0.0900, 1e-5)
    approx(z[60, 0], -4.213, 1e-5)
    xy = orc.all_distances(xyz, mem, prot, "XY", box)
    assert xy.shape == (6144, 61)
    approx(xy.max(), 9.190487, 1e-5)
    approx(xy.min(), 0.02607, 1e-5)
    approx(xy[0, 0], 3.747651)
    approx(xy[1240, 12], 3.7207017)
    approx(xy[12, 34], 6.2494035)
    approx(xy[6143, 60], 4.7850933)
    # fused consumer agrees with numpy on value and on Rust's tie rules (first min / last max)
    mn, imn, mx, imx, cnt = orc.all_distances_minmax(xyz, mem, prot, "XY", box, cutoff=1.0)
    assert mn == xy.min() and mx == xy.max()
    flat = xy.reshape(-1)
    assert imn[0] * 61 + imn[1] == int(np.flatnonzero(flat == mn)[0])
    assert imx[0] * 61 + imx[1] == int(np.flatnonzero(flat == mx)[-1])
    assert cnt == int((flat < np.float32(1.0)).sum())


# ---- analysis.rs:1595-1619
def test_atoms_distance_kat(example):
    xyz, box = example["xyz"], example["box"]
    n = xyz.shape[0]
    approx(orc.distance(xyz[0], xyz[1], "XYZ", box.diagonal()), 0.31040135)
    approx(orc.distance(xyz[n - 1], xyz[0], "XYZ", box.diagonal()), 6.664787, 2e-6)
    approx(orc.distance(xyz[n - 1], xyz[n - 2], "XYZ", box.diagonal()), 4.062491, 2e-6)


# ---- modifying.rs:688-734: atoms moved by +-k*L are restored by wrap; shifts are the k's
def test_atoms_wrap_restores(example):
    xyz, box = example["xyz"].copy(), example["box"]
    L = box.diagonal()
    inside = np.all((xyz > 0.01) & (xyz < L - 0.01), axis=1)
    ids = np.flatnonzero(inside)[:12]
    moved = xyz.copy()
    ks = np.array([[1, 0, 0], [-1, 0, 0], [0, 2, 0], [0, -1, 0], [0, 0, 1], [0, 0, -3], [1, 1, 0], [-1, 0, 2], [2, -2, 1],
                   [0, 1, -1], [-2, -1, -1], [1, 1, 1]], np.float32)
    moved[ids] += ks * L
    w, sh = orc.wrap(moved, np.arange(xyz.shape[0]), box)
    assert np.abs(w - xyz).max() < 1e-5
    assert np.array_equal(sh[ids], (-ks).astype(np.int8))
    untouched = np.setdiff1d(np.flatnonzero(inside), ids)
    assert np.array_equal(w[untouched].view(np.uint32), xyz[untouched].view(np.uint32))
    assert not sh[untouched].any()


# ---- make_group_whole / make_molecules_whole (modifying.rs:1079-1153), atoms_center (utility.rs:336-520)
def test_make_whole_goldens(conect):
    from conftest import mol_refs
    xyz, box, n = conect["xyz"], conect["box"], conect["xyz"].shape[0]
    moved, _ = orc.translate(xyz, np.arange(n), conect["translate"], box)
    # whole_group_expected.gro / whole_molecules_expected.gro are written with 3 decimals: half a quantum + f32 noise
    assert np.abs(orc.make_group_whole(moved, np.arange(n), box) - conect["whole_group"]).max() < 5.1e-4
    ref = mol_refs(n, conect["bonds"])
    assert (ref == 0).sum() == n - 1 and ref[-1] == orc.NO_MOL  # one bonded molecule, the last atom is free
    whole = orc.make_molecules_whole(moved, ref, box)
    assert np.abs(whole - conect["whole_molecules"]).max() < 5.1e-4
    assert np.array_equal(whole[-1].view(np.uint32), moved[-1].view(np.uint32))  # monoatomic: untouched


def test_make_group_whole_artificial():
    # modifying.rs:1079-1106: box 5^3... three atoms around the corner end up on the same side
    xyz = np.array([[1.0, 4.0, 2.0], [4.0, 1.0, 2.0], [1.0, 1.0, 2.0]], np.float32)
    w = orc.make_group_whole(xyz, [0, 1, 2], [5.0, 5.0, 5.0])
    d = w[:, None, :] - w[None, :, :]
    assert np.abs(d).max() <= 2.5 + 1e-5  # nobody is further than half a box from anybody else
    assert np.abs(((w - xyz) / 5.0) - np.rint((w - xyz) / 5.0)).max() < 1e-6  # only whole box vectors were added


@pytest.mark.parametrize("dim,a1,a2", [
    ("None", (9.497, 1.989, 7.498), (8.829, 11.186, 2.075)),
    ("X", (6.1465545, 1.989, 7.498), (5.478555, 11.186, 2.075)),
    ("Y", (9.497, 6.033055, 7.498), (8.829, 2.2167444, 2.075)),
    ("Z", (9.497, 1.989, 7.6634398), (8.829, 11.186, 2.2404397)),
    ("XY", (6.1465545, 6.033055, 7.498), (5.478555, 2.2167444, 2.075)),
])
def test_atoms_center_kats(example, dim, a1, a2):
    # utility.rs:336-450: first and last atom of example.gro after centering Protein
    c = orc.atoms_center(example["xyz"], example["Protein"], dim, example["box"])
    assert np.abs(c[0] - np.array(a1, np.float32)).max() < 2e-6, c[0]
    assert np.abs(c[-1] - np.array(a2, np.float32)).max() < 2e-6, c[-1]
    if dim != "None":
        e = orc.estimate_center(c, example["Protein"], example["box"])
        bc = example["box"].diagonal() / 2
        for k, ax in enumerate("XYZ"):
            if ax in dim:
                assert abs(e[k] - bc[k]) < 1e-4  # the group's estimate now sits at the box centre (utility.rs:362-366)


# ---- rmsd.rs:618-780 (Kabsch, synthetic).  nalgebra Matrix3::from([[..]]) lists COLUMNS.
def _cols(m):
    return np.array(m, np.float32).T


def test_kabsch_kats():
    e = np.eye(3, dtype=np.float32)
    c3 = [0.3333333] * 3
    r, t, rm = orc.kabsch(e, e, [1, 1, 1], c3, c3, 3.0)
    assert np.abs(r - np.eye(3)).max() < 1e-6 and np.abs(t).max() < 1e-6 and rm < 1e-6
    q = [[0.6666667, 1.0, 0.0], [-0.3333333, 0.0, 0.0], [0.6666667, 0.0, 1.0]]
    r, t, rm = orc.kabsch(e, q, [1, 1, 1], c3, c3, 3.0)
    assert np.abs(r - _cols([[0, -1, 0], [1, 0, 0], [0, 0, 1]])).max() < 1e-6 and rm < 1e-6
    q = [[2, 1, 1], [1, 2, 1], [1, 1, 2]]
    r, t, rm = orc.kabsch(e, q, [1, 1, 1], c3, [1.3333333] * 3, 3.0)
    assert np.abs(r - np.eye(3)).max() < 1e-6 and np.abs(t - 1.0).max() < 1e-6 and rm < 1e-6
    q = [[1.6666666, 2.0, 1.0], [0.6666666, 1.0, 1.0], [1.6666666, 1.0, 2.0]]
    r, t, rm = orc.kabsch(e, q, [1, 1, 1], c3, [1.3333333] * 3, 3.0)
    assert np.abs(r - _cols([[0, -1, 0], [1, 0, 0], [0, 0, 1]])).max() < 1e-6 and rm < 1e-6
    p = [[4.3, 2.1, -5.2], [1.4, 2.1, 3.9], [2.4, -3.3, 1.8]]
    q = [[2.2, 0.0, 4.6], [-1.4, 0.2, 0.3], [1.3, 9.9, 11.3]]
    r, t, rm = orc.kabsch(p, q, [1, 1, 1], [2.7, 0.3, 0.16666667], [0.7, 3.3666667, 5.4], 3.0)
    exp = _cols([[0.8842437, -0.10340805, -0.45543456], [0.2840647, -0.65496445, 0.70023507],
                 [-0.37070346, -0.7485511, -0.5497733]])
    assert np.abs(r - exp).max() < 1e-6
    assert np.abs(t - np.array([-2.0, 3.066666, 5.233333])).max() < 2e-6
    approx(rm, 4.471225, 1e-6)


RMSD_TRAJ = [0.23669721, 0.2634763, 0.26021627, 0.21364464, 0.22166993, 0.19383307, 0.26422343, 0.27013618, 0.26398134,
             0.23475659, 0.24208021]


# ---- rmsd.rs:796-820 (reference = example.tpr; we use example.gro, whose 3-decimal coordinates are the tpr's)
def test_rmsd_trajectory_kat(example, short_traj):
    prot, m = example["Protein"], short_traj["protein_mass"]
    for f in range(11):
        rm, _ = orc.calc_rmsd(example["xyz"], prot, example["box"], m, short_traj["frames"][f], prot, short_traj["boxes"][f])
        approx(rm, RMSD_TRAJ[f], 1e-6)


# ---- cfg1: protein.gro + short_trajectory_protein.xtc give the same vector (61-atom sub-system)
def test_rmsd_protein_subsystem(protein):
    idx = np.arange(61)
    for f in range(11):
        rm, _ = orc.calc_rmsd(protein["xyz"], idx, protein["box"], protein["mass"], protein["frames"][f], idx,
                              protein["boxes"][f])
        approx(rm, RMSD_TRAJ[f], 1e-6)


# ---- rmsd.rs:844-866: reference broken at PBC -> 0
def test_rmsd_broken_at_pbc(example, short_traj):
    prot, m, box = example["Protein"], short_traj["protein_mass"], example["box"]
    moved, _ = orc.translate(example["xyz"], np.arange(example["xyz"].shape[0]), [3.2, -2.1, -4.6], box)
    a, _ = orc.calc_rmsd(moved, prot, box, m, example["xyz"], prot, box)
    b, _ = orc.calc_rmsd(example["xyz"], prot, box, m, moved, prot, box)
    assert a < 1e-4 and b < 1e-4


# ---- rmsd.rs:952-994: golden fitted trajectory short_trajectory_fit.xtc (xtc precision 100 -> 0.01 nm quanta)
def test_rmsd_fit_golden(example, short_traj):
    prot, m = example["Protein"], short_traj["protein_mass"]
    worst = 0.0
    for f in range(11):
        rm, fitted = orc.calc_rmsd_and_fit(example["xyz"], prot, example["box"], m, short_traj["frames"][f], prot,
                                           short_traj["boxes"][f])
        approx(rm, RMSD_TRAJ[f], 1e-6)
        worst = max(worst, float(np.abs(fitted - short_traj["fit"][f]).max()))
    # half an xtc quantum (0.005 nm) + the reference's own f32 nalgebra-SVD noise: the residual between the
    # exact optimal rotation (f64 SVD of the same H) and the golden is a pure rotation of <= 3.3e-5 rad
    # (measured, DESIGN.md "Oracle pinning"), i.e. <= 2.5e-4 nm at the box edge.
    assert worst < 0.0053, worst


# ---- exact64 flavour agrees with ref32 where ref32's own drift is small
def test_exact64_agrees(example, short_traj):
    xyz, box, prot, mem = example["xyz"], example["box"], example["Protein"], example["Membrane"]
    assert np.abs(orc.get_center_x64(xyz, prot, box) - orc.get_center(xyz, prot, box)).max() < 2e-6
    assert np.abs(orc.estimate_center_x64(xyz, mem, box) - orc.estimate_center(xyz, mem, box)).max() < 2e-5
    m = short_traj["protein_mass"]
    r64, _ = orc.calc_rmsd_x64(xyz, prot, box, m, short_traj["frames"][3], prot, short_traj["boxes"][3])
    approx(r64, RMSD_TRAJ[3], 1e-6)


# ---- error semantics
def test_error_codes(example):
    xyz = example["xyz"]
    tric = np.array([[4, 0, 0], [0, 4, 0], [1, 0, 4]], np.float32)
    with pytest.raises(orc.OracleError) as e:
        orc.get_center(xyz, [0, 1], tric)
    assert e.value.code == orc.ENOTORTHO
    with pytest.raises(orc.OracleError) as e:
        orc.get_center(xyz, [], example["box"])
    assert e.value.code == orc.EEMPTY
    with pytest.raises(orc.OracleError) as e:
        orc.get_center(xyz, [0], np.zeros((3, 3), np.float32))
    assert e.value.code == orc.EZEROBOX
    assert np.isnan(orc.estimate_center(xyz, [], example["box"])).all()  # iterators.rs:1183-1185


# ---- hbonds.rs:505-587: the restated analyze_single against the reference's golden counts and bonds
@pytest.mark.parametrize("frame", [0, 11, 20])
def test_hbonds_water_goldens(frame):
    import os
    from oracle import xdrfile_ref
    import hbond_goldens as hg
    here = os.path.dirname(os.path.abspath(__file__))
    traj = xdrfile_ref.read_xtc(os.path.join(here, "golden", "xtc", "aa_membrane_peptide.xtc"))
    water = np.load(os.path.join(here, "golden", "aa_membrane_water.npz"))
    ow = water["OW"].astype(np.int64)
    donors = [(int(o), [int(o) + 1, int(o) + 2]) for o in ow]
    L = traj["box"][frame].reshape(3, 3).diagonal().copy()
    hg.check_frame(frame, orc.hbonds_single(traj["xyz"][frame], ow, donors, L, hg.MAX_DISTANCE, hg.MIN_ANGLE))


def test_guess_bonds_restatement_small():
    """identify_bonds (guess.rs:427-470) on a hand-made case: the limit is (vdw1 + vdw2) * factor, strict; atoms without a
    radius get no bonds; periodic images count"""
    L = [3.0, 3.0, 3.0]
    x = np.array([[0.1, 0.1, 0.1], [0.196, 0.1, 0.1], [0.07, 0.19, 0.1], [0.38, 0.1, 0.1], [2.95, 0.1, 0.1], [2.99, 0.1, 0.1]], np.float32)
    vdw = [0.152, 0.12, 0.12, 0.152, -1.0, 0.12]
    got = orc.guess_bonds(x, vdw, L, 0.55)
    # O(0)-H(1) 0.096 < 0.1496; O(0)-H(2) 0.0949; H(1)-H(2): 0.155 > 0.132; O(0)-O(3) 0.28 > 0.167; atom 4 has no radius;
    # atom 5 sits 0.11 nm from atom 0 and 0.12 nm from atom 2 THROUGH the box face: 0.11 < 0.1496, 0.12 < 0.132 -> bonds
    assert got.tolist() == [[0, 1], [0, 2], [0, 5], [2, 5]]
