#!/usr/bin/env python
"""Copies the reference's xtc FIXTURES (test data, not source) from /root/reference/test_files into tests/golden/xtc/
(TEST INFRASTRUCTURE; runs only in the build container).  The GPU box has no /root/reference, so the files the xtc codec
tests decode travel with the repo: the six trajectories SURVEY.md 8c names and the golden fitted trajectory of
rmsd.rs:952-994."""
import os
import shutil

REF = os.environ.get("GROAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "xtc")
FILES = ["short_trajectory_protein.xtc", "aa_membrane_peptide.xtc", "short_trajectory.xtc", "triclinic_trajectory.xtc",
         "dodecahedron_trajectory.xtc", "octahedron_trajectory.xtc", "short_trajectory_fit.xtc"]

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, "test_files", f), os.path.join(OUT, f))
        os.chmod(os.path.join(OUT, f), 0o644)
        print("copied", f, os.path.getsize(os.path.join(OUT, f)))
