import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def dequant(q, prec=100.0):
    """xtc decode: int * (1/precision) in f32 (proved bit exact by oracle/gen_golden.py)"""
    return q.astype(np.float32) * (np.float32(1.0) / np.float32(prec))


@pytest.fixture(scope="session")
def example():
    return load_golden("example")


@pytest.fixture(scope="session")
def short_traj():
    g = load_golden("short_trajectory")
    return {"frames": dequant(g["q"]), "boxes": g["box"], "fit": dequant(g["fit_q"]), "fit_boxes": g["fit_box"],
            "protein_mass": g["protein_mass"]}


@pytest.fixture(scope="session")
def protein():
    return load_golden("protein")


@pytest.fixture(scope="session")
def aa_pep():
    return load_golden("aa_membrane_peptide")


@pytest.fixture(scope="session")
def tric():
    return load_golden("triclinic")


@pytest.fixture(scope="session")
def conect():
    """conect.pdb (50 atoms, one CONECT-bonded molecule + one free atom) with the reference's make-whole goldens"""
    return load_golden("conect")


def mol_refs(n, bonds):
    """System::create_mol_references (modifying.rs:258-283): lowest index of every polyatomic molecule, per atom"""
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    bonded = np.zeros(n, bool)
    for a, b in np.asarray(bonds, dtype=np.int64).reshape(-1, 2):
        bonded[a] = bonded[b] = True
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    return np.array([find(i) if bonded[i] else 0xFFFFFFFF for i in range(n)], np.uint32)
