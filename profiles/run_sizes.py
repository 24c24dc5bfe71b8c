"""Device time of the single-pass ops against the group size (F frames per call): the small sizes show the fixed cost of a
call -- launch plus the serial per-frame finish (reduction re-read, Kabsch, image placement) -- that the large sizes hide.

    python profiles/run_sizes.py [F]
"""
import sys
import numpy as np
import torch
import groan_rs_b200 as g

F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
BOX = 34.0
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def time_op(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for n in (16384, 65536, 262144, 1048576, 4000000):
    rng = np.random.default_rng(1)
    m = rng.uniform(1.0, 100.0, n).astype(np.float32)
    s = g.System(n, masses=m, max_frames=F)
    ref = g.System(n, masses=m, max_frames=1)
    s.set_stream(stream.cuda_stream)
    idx = np.arange(n, dtype=np.uint32)
    s.group_create_from_indices("G", idx)
    ref.group_create_from_indices("G", idx)
    scale = 8.0 / 131070.0
    ref.set_frames(s.synth_blob_ref(7, scale, [BOX / 2] * 3), [BOX] * 3)
    rot = np.tile(np.eye(3, dtype=np.float32).reshape(1, 9), (F, 1))
    cen = rng.uniform(0, BOX, size=(F, 3)).astype(np.float32)
    s.synth_blob(7, 0, F, scale, 0.05 / 37837.23, rot, cen, [BOX] * 3, wrap=True)
    d_cen = torch.empty((F, 3), dtype=torch.float32, device=dev)
    d_rmsd = torch.empty((F,), dtype=torch.float32, device=dev)
    tc = time_op(lambda: s.group_get_center("G", out=d_cen))
    tr = time_op(lambda: s.calc_rmsd(ref, "G", out=d_rmsd))
    tf = time_op(lambda: s.group_center_and_rmsd(ref, "G", center_out=d_cen, rmsd_out=d_rmsd))
    fb = s.fallback_frames()
    print("n %8d F %d  centre %7.1f us  rmsd %7.1f us  centre+rmsd %7.1f us  (fallback frames %d)  rmsd[0] %.5f" %
          (n, F, tc, tr, tf, fb, float(d_rmsd[0])))
