"""Host-side logic of the frame-sharded multi-GPU path (groan_rs_b200/parallel.py), exercised on CPU with the
gloo backend and world_size 2 -- the same way the reference tests its threaded iteration by comparing the set of
visited frames with a serial run (parallel.rs:934-1150) and error propagation (parallel.rs:1695-1750)."""
import os
import socket

import numpy as np
import pytest

from groan_rs_b200.parallel import ParallelTrajData, frame_range, gather_frames, traj_iter_map_reduce


def test_frame_range_partitions_exactly():
    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = frame_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [frame_range(n, r, world)[1] - frame_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        frame_range(10, 2, 2)


class Visited(ParallelTrajData):
    def __init__(self):
        self.frames, self.rank = [], None

    def initialize(self, rank):
        self.rank = rank

    @staticmethod
    def reduce(items):
        out = Visited()
        for it in items:
            out.frames.extend(it.frames)
        return out


def test_map_reduce_serial_visits_every_selected_frame():
    def body(batch, idx, data):
        assert list(batch) == list(idx)
        data.frames.extend(int(i) for i in idx)

    d = traj_iter_map_reduce(103, lambda idx: idx, body, Visited(), batch_frames=16, start=5, end=99, step=3)
    assert d.frames == list(range(5, 99, 3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 11  # ragged: 6 + 5
        lo, hi = frame_range(n, rank, world)
        local = np.stack([np.arange(lo, hi, dtype=np.float32) * 10 + k for k in range(3)], axis=1)
        full = gather_frames(local, n)
        exp = np.stack([np.arange(n, dtype=np.float32) * 10 + k for k in range(3)], axis=1)
        ok_gather = np.array_equal(full, exp)
        t = gather_frames(torch.arange(lo, hi, dtype=torch.float32), n)
        ok_gather = ok_gather and torch.equal(t, torch.arange(n, dtype=torch.float32))

        def body(batch, idx, data):
            data.frames.extend(int(i) for i in idx)

        d = traj_iter_map_reduce(57, lambda idx: idx, body, Visited(), batch_frames=4, start=2, step=2)
        ok_mr = d.frames == list(range(2, 57, 2))

        def bad_body(batch, idx, data):
            if rank == 1:
                raise ValueError("boom on rank 1")

        try:
            traj_iter_map_reduce(20, lambda idx: idx, bad_body, Visited(), batch_frames=4)
            ok_err = False
        except (ValueError, RuntimeError) as e:
            ok_err = "boom" in str(e)
        q.put((rank, ok_gather, ok_mr, ok_err))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r
