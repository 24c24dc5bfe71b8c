"""Small driver for profiling the cell-grid pair search: 200 000 atoms against all 1M atoms, cutoff 1.0 nm (bench extra)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import groan_rs_b200 as g
N = 1_000_000
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
p = g.System(N, max_frames=1); p.set_stream(st.cuda_stream)
p.group_create_from_indices("Q", np.arange(500000, 700000)); p.group_create_from_indices("all1M", np.arange(N))
p.synth_uniform(20261018, 0, 1, [-2.15] * 3, [25.8] * 3, [21.5] * 3)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    c = p.group_pairs_within("Q", "all1M", 1.0)[0]
torch.cuda.synchronize()
print("ok", c)
