"""Frame-sharded trajectory iteration over the GPUs of one box: the B200 counterpart of
System::traj_iter_map_reduce (src/system/parallel.rs:208-269).

The reference runs T threads, thread n taking frames n, n+T, ... (parallel.rs:425-448), each with a clone
of the System and of the user's Data, and folds the per-thread Data with ParallelTrajData::reduce
(parallel.rs:31-49).  Here one process drives one GPU (torchrun), rank r takes the CONTIGUOUS frame range
[r*F/R, (r+1)*F/R) so that results are in global frame order after a plain concatenation, frames go to the
GPU in batches, and the only exchange is one all_gather of the small per-frame results (12 B/frame for a
centre, 4 B/frame for an RMSD) over NCCL/NVLink -- or gloo in the CPU tests.  There is no data-path
collective: frames are independent (parallel.rs:52-55).
"""
import numpy as np


def frame_range(n_frames, rank, world_size):
    """Contiguous shard [lo, hi) of rank `rank`; shard sizes differ by at most one frame."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def gather_frames(local, n_frames_total, device=None):
    """all_gather of per-frame results: `local` is [f_local, ...] for this rank's frame_range; returns the
    [n_frames_total, ...] array in global frame order on every rank.  Ragged shards are padded to the longest."""
    import torch
    dist = _dist()
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if dist is None or dist.get_world_size() == 1:
        return t if isinstance(local, torch.Tensor) else t.numpy()
    world, rank = dist.get_world_size(), dist.get_rank()
    if device is None:
        device = t.device if dist.get_backend() != "nccl" else torch.device("cuda", torch.cuda.current_device())
    sizes = [frame_range(n_frames_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if t.shape[0] != hi - lo:
        raise ValueError("rank %d holds %d frames, expected %d" % (rank, t.shape[0], hi - lo))
    if (dist.get_backend() == "nccl" and n_frames_total % world == 0 and isinstance(local, torch.Tensor) and t.is_cuda
            and t.is_contiguous()):
        # equal shards already on the exchange device: one collective straight into the result, no padding copies
        full = torch.empty((n_frames_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t)
        return full
    pad = torch.zeros((longest,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    pad[: hi - lo] = t.to(device)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    full = torch.cat([parts[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
    return full if isinstance(local, torch.Tensor) else full.cpu().numpy()


class ParallelTrajData:
    """parallel.rs:31-49: user data carried through the iteration; `reduce` folds one instance per rank."""

    def initialize(self, rank):  # parallel.rs:40
        pass

    @staticmethod
    def reduce(items):  # parallel.rs:48
        raise NotImplementedError


def traj_iter_map_reduce(n_frames, load_batch, body, init_data, batch_frames=64, start=0, end=None, step=1):
    """Rank-sharded map-reduce over a trajectory.

    load_batch(frame_indices) -> whatever `body` needs for those frames (typically it calls
    System.set_frames and returns the System); body(batch, frame_indices, data) analyses the batch and
    updates `data`; init_data is a ParallelTrajData (deep-copied per rank, initialize(rank) called on it).
    start / end / step select frames like with_range / with_step (traj_read.rs:215).  Every rank returns
    the reduced Data.  Errors raised by `body` on any rank abort all ranks (parallel.rs:453-475).
    """
    import copy
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    end = n_frames if end is None else min(end, n_frames)
    selected = np.arange(start, end, step, dtype=np.int64)
    lo, hi = frame_range(len(selected), rank, world)
    data = copy.deepcopy(init_data)
    data.initialize(rank)
    err = None
    try:
        for b in range(lo, hi, batch_frames):
            idx = selected[b: min(b + batch_frames, hi)]
            body(load_batch(idx), idx, data)
    except Exception as e:  # propagate after every rank has reached the exchange
        err = e
    if dist and world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (data if err is None else None, repr(err) if err else None))
        bad = [g[1] for g in gathered if g[1] is not None]
        if err is not None:
            raise err
        if bad:
            raise RuntimeError("traj_iter_map_reduce aborted: another rank failed with %s" % bad[0])
        return type(init_data).reduce([g[0] for g in gathered])
    if err is not None:
        raise err
    return type(init_data).reduce([data])
