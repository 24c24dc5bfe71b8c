"""Times group_get_center / calc_rmsd / group_center_and_rmsd of the bench workload (configs[4]: 4M atoms x 37 frames, device-resident)
for ONE build of the library: python profiles/exp/quad_time.py groan_rs_b200/libquad_x.so.  Tuning tool (bursts of 50 launches,
best and median of 5), not a bench line; results are checked against the separate calls and the analytic RMSD."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from groan_rs_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
    import ctypes
    probe = ctypes.CDLL(_lib.LIB_PATH)
    for name in list(_lib.SIGNATURES):
        if not hasattr(probe, name):
            _lib.SIGNATURES.pop(name)  # an older build of the library
import torch  # noqa: E402
import groan_rs_b200 as g  # noqa: E402
import bench  # noqa: E402

F = int(os.environ.get("FRAMES", "37"))
N = int(os.environ.get("NATOMS", bench.N_ATOMS))
m = bench.masses(N)
s = g.System(N, masses=m, device=0, max_frames=F)
ref = g.System(N, masses=m, device=0, max_frames=1)
idx = np.arange(N, dtype=np.uint32)
s.group_create_from_indices("G", idx)
ref.group_create_from_indices("G", idx)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
s.set_stream(stream.cuda_stream)
ref.set_frames(s.synth_blob_ref(bench.SEED, bench.BLOB_SCALE, [bench.BOX / 2] * 3), [bench.BOX] * 3)
F0 = int(os.environ.get("FRAME0", "0"))
rot, cen = bench.frame_params(F0, F)
s.synth_blob(bench.SEED, F0, F, bench.BLOB_SCALE, bench.NOISE_SCALE, rot, cen, [bench.BOX] * 3, wrap=True)
dev = torch.device("cuda", 0)
d_c = torch.empty((F, 3), dtype=torch.float32, device=dev)
d_r = torch.empty((F,), dtype=torch.float32, device=dev)
d_c2 = torch.empty((F, 3), dtype=torch.float32, device=dev)
d_r2 = torch.empty((F,), dtype=torch.float32, device=dev)


REPS = int(os.environ.get("REPS", "50"))


def burst(fn, reps=REPS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(5 if reps <= 50 else 3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b) / reps)
    return min(t), float(np.median(t))


ops = {"center": lambda: s.group_get_center("G", out=d_c), "rmsd": lambda: s.calc_rmsd(ref, "G", out=d_r),
       "fused": lambda: s.group_center_and_rmsd(ref, "G", center_out=d_c2, rmsd_out=d_r2)}
out = {k: burst(fn) for k, fn in ops.items()}
fb = s.fallback_frames()
sp = s.second_pass_frames() if hasattr(s._lib, "groan_gpu_second_pass_frames") else -1
dc = float((d_c - d_c2).abs().max())
dr = float((d_r - d_r2).abs().max())
ok = dc <= 4e-6 and dr <= 2e-6 and bool(((d_r - 0.0866).abs() < 1e-3).all())
alg = (12 * F + 16) * N * 1e-9
import subprocess
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
print("%-28s center %.4f/%.4f  rmsd %.4f/%.4f  fused %.4f/%.4f ms (best/median)  fused frac %.3f  fallback %d second %d  dC %.1e dR %.1e %s  [%s]"
      % (os.path.basename(_lib.LIB_PATH), *out["center"], *out["rmsd"], *out["fused"], alg / (out["fused"][0] * 1e-3) / 6454.9, fb, sp,
         dc, dr, "ok" if ok else "MISMATCH", clk))
