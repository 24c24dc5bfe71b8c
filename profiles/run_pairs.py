"""Small driver for profiling the all-pairs kernels (BASELINE configs[3]: 1M atoms, 2 000 x 200 000)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import groan_rs_b200 as g
N, n1, n2 = 1_000_000, 2000, 200_000
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
p = g.System(N, max_frames=2); p.set_stream(st.cuda_stream)
p.group_create_from_indices("A", np.arange(n1)); p.group_create_from_indices("B", np.arange(500000, 500000 + n2))
p.synth_uniform(20261018, 0, 2, [-2.15] * 3, [25.8] * 3, [21.5] * 3)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    r = p.group_all_distances_reduce("A", "B", g.Dimension.XYZ, cutoff=1.0)
p.synth_uniform(20261018, 0, 1, [-2.15] * 3, [25.8] * 3, [21.5] * 3)
mat = torch.empty((1, n1, n2), dtype=torch.float32, device="cuda")
for _ in range(reps):
    p.group_all_distances("A", "B", g.Dimension.XYZ, out=mat)
torch.cuda.synchronize()
print("ok", r["min"], r["argmin"], r["count"])
