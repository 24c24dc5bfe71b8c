"""Times the all-pairs fused reduction of configs[3] (2 000 x 200 000 atoms, 2 frames) alone: python profiles/exp/pairs_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import groan_rs_b200 as g
n1, n2, N = 2000, 200000, 1_000_000
p = g.System(N, device=0, max_frames=2)
p.group_create_from_indices("A", np.arange(n1))
p.group_create_from_indices("B", np.arange(500000, 500000 + n2))
p.synth_uniform(20261018, 0, 2, [-0.1 * 21.5] * 3, [1.2 * 21.5] * 3, [21.5] * 3)
dev = torch.device("cuda", 0)
red = {"min": torch.empty(2, dtype=torch.float32, device=dev), "argmin": torch.empty((2, 2), dtype=torch.int32, device=dev),
       "max": torch.empty(2, dtype=torch.float32, device=dev), "argmax": torch.empty((2, 2), dtype=torch.int32, device=dev),
       "count": torch.empty(2, dtype=torch.int64, device=dev)}
def run():
    return p.group_all_distances_reduce("A", "B", g.Dimension.XYZ, cutoff=1.0, out=red if os.environ.get("PAIRS_DEVICE_OUT", "1") == "1" else None)
for _ in range(3): out = run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b) / 10
pairs = 2 * n1 * n2
print("ms %.4f pairs/s %.3e frac_fp32 %.3f" % (t, pairs / (t * 1e-3), 18 * pairs / (t * 1e-3) / (148 * 128 * 1.965e9)))
print({k: (v.cpu().numpy() if hasattr(v, "cpu") else v).ravel()[:4] for k, v in out.items()})
