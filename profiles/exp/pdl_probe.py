"""Probe: do back-to-back quad-kernel calls complete (a) without fallback, (b) with the device-launched fallback, under PDL?"""
import sys, time
import numpy as np
import torch
import groan_rs_b200 as g

mode, group, reps = sys.argv[1], sys.argv[2], int(sys.argv[3])
n, F = 200_000, 4
L = np.array([12.0, 12.0, 12.0], np.float32)
masses = np.random.default_rng(3).uniform(1.0, 50.0, n).astype(np.float32)
s = g.System(n, masses=masses, max_frames=F)
s.set_flags({"pdl": 0, "plain": g.FLAG_NO_PDL, "hostfb": g.FLAG_HOST_FALLBACK}[mode])
s.synth_uniform(5, 0, F, [0, 0, 0], L, L)
fr = s.get_frames().copy()
fr[:, 1000:1000 + 8192] = 5.0 + 0.1 * fr[:, 1000:1000 + 8192]
s.set_frames(fr, np.tile(L, (F, 1)))
ref = g.System(n, masses=masses)
ref.set_frames(fr[0], L)
for x in (s, ref):
    x.group_create_from_indices("wide", np.arange(0, n))
    x.group_create_from_indices("narrow", np.arange(1000, 1000 + 8192))
dev = torch.device("cuda", 0)
d_c = torch.empty((F, 3), dtype=torch.float32, device=dev)
d_r = torch.empty((F,), dtype=torch.float32, device=dev)
t0 = time.time()
for rep in range(reps):
    if group in ("narrow", "both"):
        s.group_center_and_rmsd(ref, "narrow", center_out=d_c, rmsd_out=d_r)
    if group in ("wide", "both"):
        s.group_center_and_rmsd(ref, "wide", center_out=d_c, rmsd_out=d_r)
s.sync()
print(mode, group, reps, "ok %.3f s" % (time.time() - t0), "fallback frames", s.fallback_frames(), flush=True)
