"""Times the cell-grid pair search (200 000 x 1M atoms, cutoff 1.0 nm) in count-only and in store mode: python profiles/exp/cells_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import groan_rs_b200 as g
N = 1_000_000
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
p = g.System(N, max_frames=1); p.set_stream(st.cuda_stream)
p.group_create_from_indices("Q", np.arange(500000, 700000)); p.group_create_from_indices("all1M", np.arange(N))
p.synth_uniform(20261018, 0, 1, [-2.15] * 3, [25.8] * 3, [21.5] * 3)
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
t, out = timed(lambda: p.group_pairs_within("Q", "all1M", 1.0))
print("count only: %.3f ms, %d pairs" % (t, int(out[0][0])))
cap = 120_000_000
dev = torch.device("cuda", 0)
d_pairs = torch.empty((1, cap, 2), dtype=torch.int32, device=dev)
d_dist = torch.empty((1, cap), dtype=torch.float32, device=dev)
t, out = timed(lambda: p.group_pairs_within("Q", "all1M", 1.0, pairs_out=d_pairs), reps=3)
print("store pairs (device lists): %.3f ms, %d pairs" % (t, int(out[0][0])))
t, out = timed(lambda: p.group_pairs_within("Q", "all1M", 1.0, pairs_out=d_pairs, dist_out=d_dist), reps=3)
print("store pairs + distances: %.3f ms" % t)
