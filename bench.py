#!/usr/bin/env python
"""bench.py -- frames/s of the PBC geometry hot path (group_get_center + calc_rmsd per frame) on B200.

Workload (BASELINE.json configs[4], orthogonal box -- the reference rejects triclinic boxes for this path,
SURVEY.md section 0.1): synthetic 4 000 000-atom system, box 34 nm, a compact rigid blob (extent < half box)
that is randomly rotated, translated across the periodic boundary and perturbed by 0.05 nm noise per frame, so
the RMSD to the reference structure is ~0.087 nm analytically.  A "step" is one batch of F frames (48 MB each;
the batch is larger than the 126 MB L2) through System.group_get_center + System.calc_rmsd.

One JSON line on stdout (rank 0).  `value` times device-resident frames with CUDA events; `e2e` pushes the same
batch from pinned HOST memory through the public API every step (H2D + kernels + D2H of the results).
`--impl reference` times the restated groan_rs CPU path (oracle/, the reference itself is Rust and cannot be
built in this image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ATOMS = 4_000_000
BOX = 34.0
SEED = 20261018
BLOB_SCALE = 5.5 / 131070.0        # blob support +-5.5 nm per axis: extent (rotated) stays below half the box (17 nm)
NOISE_SCALE = 0.05 / 37837.23      # Irwin-Hall(4) of 16-bit fields has sigma 37837.23 -> 0.05 nm per axis
METRIC = "frames/s (group_get_center + calc_rmsd per frame, 4M-atom group)"


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def frame_params(frame0, n_frames):
    """per-frame random rigid motion: rotation (unit quaternion) and a centre anywhere in the box"""
    rot = np.empty((n_frames, 9), np.float32)
    cen = np.empty((n_frames, 3), np.float32)
    for k in range(n_frames):
        rng = np.random.default_rng([SEED, frame0 + k])
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        rot[k] = np.array([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                           2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                           2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], np.float32)
        cen[k] = rng.uniform(0.0, BOX, size=3).astype(np.float32)
    return rot, cen


def masses(n):
    return np.random.default_rng(SEED + 1).uniform(1.0, 100.0, n).astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin, self.t_end = None, None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.perf_counter()])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.summary()

    def summary(self):
        # the sampler is started before the warm-up (so that nvidia-smi's start-up does not fall into the timed region);
        # only the rows read between mark_begin() and mark_end() (+ one polling interval) count
        rows = [r for r in self.rows if len(r) >= 8]
        if self.t_begin is not None and self.t_end is not None:
            inside = [r for r in rows if self.t_begin <= r[-1] <= self.t_end + 0.03]
            rows = inside or rows
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(frames_per_step):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the last `ncu --set full`
    capture of this workload (profiles/r2_dominant_kernel.json); None if the capture was made with another batch size"""
    p = os.path.join(ROOT, "profiles", "r2_dominant_kernel.json")
    try:
        d = json.load(open(p))
        if d.get("frames_per_step") == frames_per_step:
            return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(steps, warmup, frames_per_step, threads=None, sample_frames=None):
    """Restated groan_rs CPU trajectory path (oracle/groan_oracle.c: orc_baseline_traj_cyclic, following
    parallel.rs:208-269,425-448): T threads, each with its own clone of the System (AoS 240-byte atom records), the frames of
    the WHOLE run interleaved among them (thread t takes frames t, t + T, ...), f32 sequential sums.
    A step is `frames_per_step` frames like in the GPU arm; `sample_frames` distinct frames are held in memory and visited
    cyclically (37 x 48 MB; a run of 20 steps would otherwise need 35 GB)."""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    T = threads or max(1, min(cores, 32))
    Fs = sample_frames or frames_per_step
    m = masses(N_ATOMS)
    idx = np.arange(N_ATOMS, dtype=np.uint32)
    L = np.array([BOX, BOX, BOX], np.float32)
    ref_xyz = orc.synth_blob_ref(N_ATOMS, SEED, BLOB_SCALE, [BOX / 2] * 3)
    rot, cen = frame_params(0, Fs)
    frames = np.empty((Fs, N_ATOMS, 3), np.float32)
    for f in range(Fs):
        frames[f] = orc.synth_blob_frame(N_ATOMS, SEED, f, BLOB_SCALE, NOISE_SCALE, rot[f], cen[f], L, wrap=True)
    boxes = np.tile(L, (Fs, 1))
    if warmup > 0:
        orc.baseline_traj(frames, boxes, idx, m, ref_xyz, L, ops=1 | 2, n_threads=T, total_frames=warmup * frames_per_step)
    total = max(1, steps) * frames_per_step
    sec, _, _ = orc.baseline_traj(frames, boxes, idx, m, ref_xyz, L, ops=1 | 2, n_threads=T, total_frames=total)
    return {"value": total / sec, "unit": "frames/s", "cores": T, "kind": "port",
            "sample": "%d frames (= %d steps x %d; %d distinct frames visited cyclically) x %d atoms (group_get_center + calc_rmsd), "
                      "%d threads, frames interleaved among the threads like traj_iter_map_reduce; restated groan_rs CPU path (oracle/), "
                      "generation excluded" % (total, max(1, steps), frames_per_step, Fs, N_ATOMS, T),
            "ms_per_step": sec * 1e3 / max(1, steps), "frames": frames_per_step}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = max(args.warmup, 0)
    r = cpu_reference_run(max(1, args.steps), W, args.frames, sample_frames=args.cpu_frames)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded generator shared with the GPU arm)",
            "config": workload_config(r["frames"]),
            "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "warmup_steps_run": W, "gpu_launches": 0}
    emit(line)


def workload_config(frames_per_step):
    return {"workload": "configs[4] synthetic 4M-atom box 34 nm (orthogonal: the reference rejects triclinic for this path), "
                        "group = all 4M atoms, per frame group_get_center + calc_rmsd vs reference structure",
            "n_atoms": N_ATOMS, "group_atoms": N_ATOMS, "frames_per_step": frames_per_step, "box_nm": BOX,
            "batching": "one call per step over a batch of %d frames (%d MB resident in HBM)" % (frames_per_step,
                                                                                                 frames_per_step * N_ATOMS * 12 // 1000000),
            "l2_policy": "inputs larger than L2: %d MB per step vs 126 MB L2" % (frames_per_step * N_ATOMS * 12 // 1000000)}


class CentreRows:
    """ParallelTrajData of the bench's traj_iter_map_reduce check (module level: it travels between ranks by pickle)"""

    def __init__(self):
        self.rows = {}

    def initialize(self, rank):  # parallel.rs:40
        pass

    @staticmethod
    def reduce(items):  # parallel.rs:48
        out = CentreRows()
        for it in items:
            out.rows.update(it.rows)
        return out


def pin_to_gpu_numa(local):
    """Run this rank's host side (threads, first-touch of its pinned buffers) on the NUMA node its GPU hangs off.
    Returns what was found for the JSON line; a VM that exposes no topology reports node -1 and nothing is changed."""
    info = {"node": None, "cpus": None}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        info["pci"] = bus
        info["node"] = node
        if node >= 0:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"] = len(allowed)
    except Exception as e:  # no sysfs topology (containers, VMs): leave the affinity alone
        info["error"] = type(e).__name__
    return info


def check_frames_against_oracle(ref_xyz, m, rot, cen, frame0, which, got_center, got_rmsd):
    """exact64 oracle (oracle/groan_oracle.c: orc_get_center_x64, orc_calc_rmsd_x64) on frames `which` of the batch;
    asserts the north-star tolerances and returns the largest deviations for the JSON line"""
    from oracle import oracle as orc
    idx = np.arange(N_ATOMS, dtype=np.uint32)
    L = np.array([BOX, BOX, BOX], np.float32)
    dc, dr = 0.0, 0.0
    for f in which:
        fr = orc.synth_blob_frame(N_ATOMS, SEED, frame0 + f, BLOB_SCALE, NOISE_SCALE, rot[f], cen[f], L, wrap=True)
        c64 = orc.get_center_x64(fr, idx, L)
        r64, _ = orc.calc_rmsd_x64(ref_xyz, idx, L, m, fr, idx, L)
        dc = max(dc, float(np.abs(got_center[f] - c64).max()))
        dr = max(dr, abs(float(got_rmsd[f]) - float(r64)))
    assert dc <= 1e-5 and dr <= 1e-4, (dc, dr)
    return {"frames_checked": list(which), "max_center_err_nm": dc, "max_rmsd_err_nm": dr, "oracle": "exact64 (oracle/)",
            "tolerance_nm": {"center": 1e-5, "rmsd": 1e-4}}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    import groan_rs_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F, K, W = args.frames, args.steps, max(3, args.warmup)

    m = masses(N_ATOMS)
    s = g.System(N_ATOMS, masses=m, device=local, max_frames=F)
    ref = g.System(N_ATOMS, masses=m, device=local, max_frames=1)
    idx = np.arange(N_ATOMS, dtype=np.uint32)
    s.group_create_from_indices("G", idx)
    ref.group_create_from_indices("G", idx)
    # a non-default torch stream: torch.cuda.Event then times exactly the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    s.set_stream(stream.cuda_stream)
    ref_xyz = s.synth_blob_ref(SEED, BLOB_SCALE, [BOX / 2] * 3)
    ref.set_frames(ref_xyz, [BOX] * 3)

    # this rank's frames: contiguous shard of the global trajectory (parallel.py frame_range); weak scaling
    frame0 = rank * F
    rot, cen = frame_params(frame0, F)
    s.synth_blob(SEED, frame0, F, BLOB_SCALE, NOISE_SCALE, rot, cen, [BOX] * 3, wrap=True)
    # per-frame results of every step of the run stay on the device, in trajectory order; ONE exchange at the end of the
    # timed region gathers them on all ranks (the only exchange of the path, 16 B per frame); the ranks never
    # wait for each other inside the loop.
    n_slots = max(K, W, 1)
    cen_all = torch.empty((n_slots * F, 3), dtype=torch.float32, device=dev)
    rmsd_all = torch.empty((n_slots * F,), dtype=torch.float32, device=dev)
    d_cen, d_rmsd = cen_all[:F], rmsd_all[:F]  # slot 0, also used by the per-op timings below
    step_no = [0]

    def step():
        # group_get_center + calc_rmsd of the same group: one read of every frame (groan_gpu_center_rmsd)
        i = step_no[0] % n_slots
        s.group_center_and_rmsd(ref, "G", center_out=cen_all[i * F:(i + 1) * F], rmsd_out=rmsd_all[i * F:(i + 1) * F])
        step_no[0] += 1

    gathered = [None, None]

    def drain():
        # the path's only exchange, through the product's own gather: groan_rs_b200.parallel.gather_frames
        step_no[0] = 0
        if world > 1:
            gathered[0] = g.gather_frames(cen_all, world * n_slots * F)
            gathered[1] = g.gather_frames(rmsd_all, world * n_slots * F)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        step()
    drain()
    barrier()
    # correctness guard inside the bench (outside the timed region): the first and the last frame of this rank's batch are
    # generated again on the host (oracle/ generator, bit-identical to the device's: tests/test_gpu_parity.py
    # test_synth_generators_match_oracle) and evaluated by the exact64 oracle -- centre within 1e-5 nm, RMSD within 1e-4 nm
    # (the north-star tolerances); every other frame must at least sit at the generator's analytic RMSD
    r0, c0 = d_rmsd.cpu().numpy(), d_cen.cpu().numpy()
    assert np.all(np.abs(r0 - 0.0866) < 1e-3), r0
    parity = check_frames_against_oracle(ref_xyz, m, rot, cen, frame0, sorted({0, F - 1}), c0, r0)
    # ... and the timed path must be the single-pass kernels, not their reference-order fallback
    fallback = {}
    s.group_get_center("G", out=d_cen)
    fallback["group_get_center"] = s.fallback_frames()
    s.calc_rmsd(ref, "G", out=d_rmsd)
    fallback["calc_rmsd"] = s.fallback_frames()
    s.group_center_and_rmsd(ref, "G", center_out=d_cen, rmsd_out=d_rmsd)
    fallback["group_center_and_rmsd"] = s.fallback_frames()
    second_pass = s.second_pass_frames()  # frames of the batch whose centre went through the sine-sum pass (second tier)

    l0 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(K):
        step()
    drain()  # every batch's results have been gathered before the clock stops
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    launches = s.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * F * K / (ms * 1e-3)

    # per-op device times (same resident batch, CUDA events on the launching stream)
    def time_op(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    peak, peak_src = measured_peak()
    reps = max(3, min(K, 50))
    t_center = time_op(lambda: s.group_get_center("G", out=d_cen), reps)
    t_rmsd = time_op(lambda: s.calc_rmsd(ref, "G", out=d_rmsd), reps)
    t_fused = time_op(lambda: s.group_center_and_rmsd(ref, "G", center_out=d_cen, rmsd_out=d_rmsd), reps)
    gb = 1e-9
    ops = {
        "group_get_center": {"ms": t_center, "alg_bytes": 12 * N_ATOMS * F, "gbs": 12 * N_ATOMS * F * gb / (t_center * 1e-3)},
        # compulsory HBM bytes: every frame once (12 B/atom) + the reference (pc.xyz, w: 16 B/atom) once per launch;
        # its re-reads by the other frames of the batch are served by the 126 MB L2 (evict-last)
        "calc_rmsd": {"ms": t_rmsd, "alg_bytes": (12 * F + 16) * N_ATOMS, "gbs": (12 * F + 16) * N_ATOMS * gb / (t_rmsd * 1e-3)},
        "group_center_and_rmsd": {"ms": t_fused, "alg_bytes": (12 * F + 16) * N_ATOMS,
                                  "gbs": (12 * F + 16) * N_ATOMS * gb / (t_fused * 1e-3)},
    }
    dom = "group_center_and_rmsd"  # the kernel the timed step runs (k_rmsd_quad<SAME_MASS, CENTER = 1>, kernels_quad.cuh)
    # a step IS one launch of that kernel (gpu_launches == steps), so its average launch duration over the timed region is
    # ms / K (at N > 1 that includes waiting for the slowest rank); the per-op figures below are short bursts of 50 launches
    # (since round 2 a step is that kernel plus the host-launched second-tier pass, whose CTAs exit at once unless a frame's
    # mean sits within ~0.1 nm of a box face: gpu_launches == 2 x steps; ms / K charges both to the dominant kernel)
    achieved = ops[dom]["alg_bytes"] * gb / (ms / K * 1e-3) if launches in (K, 2 * K) else ops[dom]["gbs"]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(F), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ops[dom]["alg_bytes"],
                "timing": "algorithmic bytes per launch / (timed region / launches), CUDA events on the launching stream",
                "achieved_burst": ops[dom]["gbs"], "frac_burst": ops[dom]["gbs"] / peak, "ops": ops,
                "fallback_frames": fallback, "second_pass_frames": second_pass,
                "launches_per_step": launches / max(K, 1)}

    # ---- end to end through the public API with HOST buffers (pinned): H2D + decode + kernels + D2H every step.
    # Four feeds of the SAME frames (rounded to the xtc lattice for the last three):
    #   f32        what read_xtc hands out (12 B/atom)                                 -> System.set_frames
    #   int16      the decoder's lattice integers + a per-frame origin (6 B/atom)      -> System.set_frames_quantized
    #   xtc        the FILE'S BYTES as they lie on disk, decoded on the GPU            -> System.set_frames_xtc
    #   xtc_host   the file's bytes decoded by the host thread pool inside the step    -> XtcFile.decode + set_frames_quantized
    # The headline `value` is the xtc feed: its host buffer is the trajectory file itself, nothing is prepared on the host.
    e2e = None
    if not args.no_e2e:
        numa = pin_to_gpu_numa(local)
        h_in = [torch.empty((F, N_ATOMS, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        s.get_frames(out=h_in[0])
        s.sync()
        t0 = time.perf_counter()
        h_in[1].copy_(h_in[0])
        host_memcpy_gbs = 2 * F * N_ATOMS * 12 * 1e-9 / (time.perf_counter() - t0)  # read + write, one thread
        # the platform's ceiling for any host-fed path: plain pinned -> device copies of the same buffer, all ranks at once,
        # nothing else running (at N = 8 the ranks share the host's memory system and PCIe root complexes)
        d_raw = torch.empty((F * N_ATOMS * 3,), dtype=torch.float32, device=dev)
        d_raw.copy_(h_in[0].view(-1), non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            d_raw.copy_(h_in[0].view(-1), non_blocking=True)
        torch.cuda.synchronize()
        dt_raw = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_raw], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_raw = float(t.item())
        h2d_raw_gbs = 4 * F * N_ATOMS * 12 * 1e-9 / dt_raw
        del d_raw
        h_cen = torch.empty((F, 3), dtype=torch.float32).pin_memory()
        h_rmsd = torch.empty((F,), dtype=torch.float32).pin_memory()
        boxes = np.tile(np.array([BOX, BOX, BOX], np.float32), (F, 1))
        Ke = max(3, min(K, 50))  # 8-30 ms per step: bounded so that the default run stays short

        def run_feed(push, steps):
            def one(k):
                push(k)
                s.group_center_and_rmsd(ref, "G", center_out=h_cen, rmsd_out=h_rmsd)
            for k in range(2):
                one(k)
            barrier()
            t0 = time.perf_counter()
            for k in range(steps):
                one(k)
            s.sync()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        feeds = {}
        dt = run_feed(lambda k: s.set_frames(h_in[k & 1], boxes), Ke)
        assert np.abs(h_rmsd.numpy() - r0).max() <= 2e-6 and np.abs(h_cen.numpy() - c0).max() <= 4e-6  # same frames, same results
        feeds["f32"] = {"value": world * F * Ke / dt, "h2d_bytes_per_step": F * N_ATOMS * 12 + F * 36, "ms_per_step": dt * 1e3 / Ke,
                        "steps": Ke, "h2d_gbs_per_gpu": (F * N_ATOMS * 12) * 1e-9 / (dt / Ke)}

        # the trajectory file: the batch written as xtc (precision 1000) by the product's encoder -- byte-identical to the
        # reference's write_xtc (tests/test_xtc_codec.py) -- held in pinned memory like a file read into a pinned buffer
        prec = 1000.0
        stream = g.xtc.encode(xyz=h_in[0].numpy(), boxes=boxes, precision=prec)
        h_x = [torch.empty(stream.size, dtype=torch.uint8).pin_memory() for _ in range(2)]
        for t_ in h_x:
            t_.numpy()[:] = stream
        xf = [g.xtc.XtcFile(t_) for t_ in h_x]
        del stream
        assert xf[0].n_frames == F and xf[0].n_atoms == N_ATOMS
        dt = run_feed(lambda k: s.set_frames_xtc(xf[k & 1]), Ke)
        assert s.xtc_bad_frames() == 0
        cx, rx = h_cen.numpy().copy(), h_rmsd.numpy().copy()
        # parity of this feed: frame 0 as the REFERENCE's reader semantics decode it (host decoder, bit-identical to read_xtc)
        # through the exact64 oracle
        from oracle import oracle as orc
        fr0 = xf[0].decode(first=0, count=1)["xyz"][0]
        idx_all = np.arange(N_ATOMS, dtype=np.uint32)
        Lb = np.array([BOX] * 3, np.float32)
        c64 = orc.get_center_x64(fr0, idx_all, Lb)
        r64, _ = orc.calc_rmsd_x64(ref_xyz, idx_all, Lb, m, fr0, idx_all, Lb)
        assert np.abs(cx[0] - c64).max() <= 1e-5 and abs(float(rx[0]) - float(r64)) <= 1e-4, (cx[0], c64, rx[0], r64)
        xtc_bytes = int(xf[0].offsets[-1])
        feeds["xtc"] = {"value": world * F * Ke / dt, "h2d_bytes_per_step": xtc_bytes + F * 56, "ms_per_step": dt * 1e3 / Ke, "steps": Ke,
                        "h2d_gbs_per_gpu": xtc_bytes * 1e-9 / (dt / Ke), "bytes_per_atom": xtc_bytes / (F * N_ATOMS),
                        "parity_frame0_vs_exact64": {"center_err_nm": float(np.abs(cx[0] - c64).max()),
                                                     "rmsd_err_nm": abs(float(rx[0]) - float(r64))},
                        "note": "host buffer = the xtc file's bytes (precision 1000); uploaded as they are, decoded on the GPU"}

        h_q = [torch.empty((F, N_ATOMS, 3), dtype=torch.int16).pin_memory() for _ in range(2)]
        origin = [np.zeros((F, 3), np.int32) for _ in range(2)]
        xf[0].decode(want="q16", out=h_q[0].numpy(), origin_out=origin[0])
        h_q[1].copy_(h_q[0])
        origin[1][:] = origin[0]
        dt = run_feed(lambda k: s.set_frames_quantized(h_q[k & 1], prec, boxes, origin=origin[k & 1]), Ke)
        assert np.array_equal(h_rmsd.numpy(), rx) and np.array_equal(h_cen.numpy(), cx)  # the same floats reach the kernels
        feeds["int16"] = {"value": world * F * Ke / dt, "h2d_bytes_per_step": F * N_ATOMS * 6 + F * 48, "ms_per_step": dt * 1e3 / Ke,
                          "steps": Ke, "h2d_gbs_per_gpu": (F * N_ATOMS * 6) * 1e-9 / (dt / Ke),
                          "note": "the xtc decoder's lattice integers (int16 + per-frame origin), floats rebuilt on the device"}
        if world == 1:
            # the host decoder inside the step: what a host-side reader sustains on this box's cores
            nthr = g.xtc.default_threads()

            def push_host_decoded(k):
                xf[k & 1].decode(want="q16", out=h_q[k & 1].numpy(), origin_out=origin[k & 1], n_threads=nthr)
                s.set_frames_quantized(h_q[k & 1], prec, boxes, origin=origin[k & 1])

            dt = run_feed(push_host_decoded, 3)
            assert np.array_equal(h_rmsd.numpy(), rx)
            feeds["xtc_host_decode"] = {"value": F * 3 / dt, "h2d_bytes_per_step": F * N_ATOMS * 6 + F * 48, "ms_per_step": dt * 1e3 / 3,
                                        "steps": 3, "host_threads": nthr,
                                        "note": "xtc bytes -> int16 lattice by the host thread pool (groan_xtc_decode) inside the step"}
        # partial frames (GroupXtcReader semantics, molly_xtc.rs:441-462): the analysis needs a 400 000-atom group (every 10th
        # atom, SURVEY 8d cfg5 "G = 400 000 scattered"), so only those atoms are read and uploaded -- 10x fewer PCIe bytes
        sel = np.arange(0, N_ATOMS, 10, dtype=np.uint32)
        for sysm in (s, ref):
            sysm.group_create_from_indices("S10", sel)
        s.set_frames(h_in[0], boxes)
        want_c, want_r = s.group_center_and_rmsd(ref, "S10")  # the same frames through the full upload
        h_sel = [torch.from_numpy(np.ascontiguousarray(h_in[0].numpy()[:, sel, :])).pin_memory() for _ in range(2)]

        def run_group_feed(steps):
            def one(k):
                s.set_group_frames(h_sel[k & 1], sel, boxes)
                s.group_center_and_rmsd(ref, "S10", center_out=h_cen, rmsd_out=h_rmsd)
            for k in range(2):
                one(k)
            barrier()
            t0 = time.perf_counter()
            for k in range(steps):
                one(k)
            s.sync()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt

        dt = run_group_feed(Ke)
        assert np.array_equal(h_rmsd.numpy(), want_r) and np.array_equal(h_cen.numpy(), want_c)
        feeds["group_f32"] = {"value": world * F * Ke / dt, "h2d_bytes_per_step": F * int(sel.size) * 12 + F * 36 + int(sel.size) * 4,
                              "ms_per_step": dt * 1e3 / Ke, "steps": Ke, "group_atoms": int(sel.size),
                              "h2d_gbs_per_gpu": F * int(sel.size) * 12 * 1e-9 / (dt / Ke),
                              "note": "only the group's atoms are uploaded (System.set_group_frames) and scattered into the resident frames; "
                                      "centre + RMSD of that 400 000-atom index-list group"}
        head = feeds["xtc"]
        e2e = {"value": head["value"], "unit": "frames/s", "h2d_bytes_per_step": head["h2d_bytes_per_step"], "d2h_bytes_per_step": F * 16,
               "ms_per_step": head["ms_per_step"], "steps": Ke, "feed": "xtc (file bytes, GPU decode)",
               "f32_equivalent_h2d_bytes_per_step": F * N_ATOMS * 12 + F * 36, "timing": "host wall clock around the steps, sync both sides",
               "feeds": feeds, "host_memcpy_gbs_one_thread": host_memcpy_gbs, "numa": numa,
               "h2d_raw_gbs_per_gpu": h2d_raw_gbs, "h2d_raw_gbs_all_gpus": h2d_raw_gbs * world,
               "h2d_raw_note": "plain cudaMemcpyAsync of the f32 batch from pinned memory on every rank at once, slowest rank: the ceiling "
                               "of any host-fed feed on this box at this N"}
        s.set_frames(h_in[0], boxes)  # back to the f32 batch for whatever follows
        del h_q, h_x, xf

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:  # secondary single-GPU numbers: not while other ranks would wait
        extras = run_extras(torch, g, local, peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_reference_run(3, 0, F, sample_frames=args.cpu_frames)  # ~15 s of CPU work: 3 steps of F frames
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    # the product's rank-sharded map-reduce (groan_rs_b200.parallel.traj_iter_map_reduce = System::traj_iter_map_reduce,
    # parallel.rs:208-269) over this launch's process group, outside the timed region: a small trajectory, every rank
    # analyses its contiguous shard in batches, the Data objects are reduced on every rank and must equal what one rank
    # computes for all frames
    parallel_check = None
    if world > 1:
        nP, FP = 65_536, 4 * world + 3  # ragged shards on purpose
        small = g.System(nP, device=local, max_frames=FP)
        small.set_stream(torch.cuda.current_stream().cuda_stream)
        small.group_create_from_indices("G", np.arange(16, nP - 5, dtype=np.uint32))
        rotP, cenP = frame_params(10_000, FP)
        cenP = (cenP * (12.0 / BOX)).astype(np.float32)

        def load(idx):
            small.synth_blob(SEED, int(idx[0]), len(idx), 1.5 / 131070.0, 0.02 / 37837.23, rotP[idx], cenP[idx], [12.0] * 3, wrap=True)
            return small

        def body(sysm, idx, data):
            c = sysm.group_get_center("G")
            for k, f in enumerate(idx):
                data.rows[int(f)] = c[k].copy()

        red = g.traj_iter_map_reduce(FP, load, body, CentreRows(), batch_frames=3)
        small.synth_blob(SEED, 0, FP, 1.5 / 131070.0, 0.02 / 37837.23, rotP, cenP, [12.0] * 3, wrap=True)
        direct = small.group_get_center("G")
        # a frame generated inside a batch that starts elsewhere is the same frame: the generator is counter-based per frame
        # (different batch sizes use different grids: the partial sums are folded in another order, hence a tolerance)
        okp = len(red.rows) == FP and all(np.abs(red.rows[f] - direct[f]).max() <= 4e-6 for f in range(FP))
        parallel_check = {"frames": FP, "ranks": world, "ok": bool(okp), "what": "traj_iter_map_reduce over NCCL == one rank over all frames"}
        assert okp
        small.close()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic: seeded rigid blob + noise generated on the device, batch resident in HBM and reused every step",
                "config": workload_config(F), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "parity": parity, "parallel_check": parallel_check, "extras": extras}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_extras(torch, g, local, peak):
    """secondary numbers of the same path: atoms_wrap (HBM) and all-pairs distances (configs[3])"""
    out = {}
    dev = torch.device("cuda", local)

    def time_op(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # atoms_wrap on 4M atoms x 8 frames: 24 B per atom per frame
    F = 8
    assert torch.cuda.current_stream().cuda_stream != 0
    w = g.System(N_ATOMS, device=local, max_frames=F)
    w.set_stream(torch.cuda.current_stream().cuda_stream)
    w.synth_uniform(SEED, 0, F, [-0.1 * BOX] * 3, [1.2 * BOX] * 3, [BOX] * 3)
    t = time_op(lambda: w.atoms_wrap())
    out["atoms_wrap"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "gbs": 24 * N_ATOMS * F * 1e-9 / (t * 1e-3),
                         "frac_of_hbm_peak": 24 * N_ATOMS * F * 1e-9 / (t * 1e-3) / peak}
    w.close()
    # configs[2] at scale: the triclinic EXTENSION (parity unpinned in the reference) -- atoms_wrap in the
    # triclinic variant of the configs[4] box (v2x = 8.5, v3x = 8.5, v3y = -8.5), 24 B per atom per frame
    tri = [BOX, 0, 0, 8.5, BOX, 0, 8.5, -8.5, BOX]
    wt = g.System(N_ATOMS, device=local, max_frames=F, triclinic=True)
    wt.set_stream(torch.cuda.current_stream().cuda_stream)
    wt.synth_uniform(SEED, 0, F, [-0.1 * BOX] * 3, [1.2 * BOX] * 3, tri)
    t = time_op(lambda: wt.atoms_wrap())
    out["atoms_wrap_triclinic"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "gbs": 24 * N_ATOMS * F * 1e-9 / (t * 1e-3),
                                   "frac_of_hbm_peak": 24 * N_ATOMS * F * 1e-9 / (t * 1e-3) / peak}
    wt.close()
    # calc_rmsd_and_fit (all 4M atoms rewritten per frame: 12 B read + 24 B fit traffic per atom) and a scattered group
    # (every 5th atom through an index list: the gather path), both on the blob workload
    m = masses(N_ATOMS)
    b = g.System(N_ATOMS, masses=m, device=local, max_frames=F)
    r = g.System(N_ATOMS, masses=m, device=local, max_frames=1)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    every5 = np.arange(0, N_ATOMS, 5, dtype=np.uint32)
    for sysm in (b, r):
        sysm.group_create_from_indices("G", np.arange(N_ATOMS, dtype=np.uint32))
        sysm.group_create_from_indices("S", every5)
    r.set_frames(b.synth_blob_ref(SEED, BLOB_SCALE, [BOX / 2] * 3), [BOX] * 3)
    rot, cen = frame_params(0, F)
    b.synth_blob(SEED, 0, F, BLOB_SCALE, NOISE_SCALE, rot, cen, [BOX] * 3, wrap=True)
    d_r = torch.empty((F,), dtype=torch.float32, device=dev)
    d_c = torch.empty((F, 3), dtype=torch.float32, device=dev)
    t = time_op(lambda: b.group_center_and_rmsd(r, "S", center_out=d_c, rmsd_out=d_r))
    nS = len(every5)
    out["scattered_group_center_and_rmsd"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "group_atoms": nS,
                                              "fallback_frames": b.fallback_frames(),
                                              "alg_gbs": (16 * F + 16) * nS * 1e-9 / (t * 1e-3)}
    t = time_op(lambda: b.calc_rmsd_and_fit(r, "G", out=d_r), reps=3)  # each call fits the already fitted frames again
    out["calc_rmsd_and_fit"] = {"ms": t, "frames_per_s": F / (t * 1e-3),
                                "gbs": ((12 + 24) * F + 16) * N_ATOMS * 1e-9 / (t * 1e-3),
                                "frac_of_hbm_peak": ((12 + 24) * F + 16) * N_ATOMS * 1e-9 / (t * 1e-3) / peak}
    # the cliff behind the single pass: the same batch through the reference-order passes only
    # (GROAN_FLAG_EXACT_ONLY) -- what a frame costs when the single pass cannot certify it
    b.set_flags(g.FLAG_EXACT_ONLY)
    t = time_op(lambda: b.group_get_center("G", out=d_c), reps=3)
    out["exact_only_group_get_center"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "gbs": 12 * N_ATOMS * F * 1e-9 / (t * 1e-3),
                                          "frac_of_hbm_peak": 12 * N_ATOMS * F * 1e-9 / (t * 1e-3) / peak}
    t = time_op(lambda: b.calc_rmsd(r, "G", out=d_r), reps=3)
    out["exact_only_calc_rmsd"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "gbs": (12 * F + 16) * N_ATOMS * 1e-9 / (t * 1e-3),
                                   "frac_of_hbm_peak": (12 * F + 16) * N_ATOMS * 1e-9 / (t * 1e-3) / peak}
    b.set_flags(0)
    # a box-spanning slab (a membrane: uniform in x and y, 4 nm thick in z): never compact in x / y, so every frame takes the
    # path behind the single pass
    b.synth_uniform(SEED, 0, F, [0.0, 0.0, 15.0], [BOX, BOX, 4.0], [BOX] * 3)
    r.set_frames(b.get_frames()[0], [BOX] * 3)
    t = time_op(lambda: b.group_get_center("G", out=d_c), reps=3)
    out["slab_group_get_center"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "fallback_frames": b.fallback_frames(),
                                    "frac_of_hbm_peak": 12 * N_ATOMS * F * 1e-9 / (t * 1e-3) / peak}
    t = time_op(lambda: b.calc_rmsd(r, "G", out=d_r), reps=3)
    out["slab_calc_rmsd"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "fallback_frames": b.fallback_frames(),
                             "frac_of_hbm_peak": (12 * F + 16) * N_ATOMS * 1e-9 / (t * 1e-3) / peak}
    b.close()
    r.close()
    # configs[4] names a TRICLINIC box: the extension of the centre / RMSD path (parity unpinned: the reference
    # returns NotOrthogonal), on the box variant of SURVEY 8d cfg5 (v2x = 8.5, v3x = 8.5, v3y = -8.5); gather kernels
    bt = g.System(N_ATOMS, masses=m, device=local, max_frames=F, triclinic=True)
    rt = g.System(N_ATOMS, masses=m, device=local, max_frames=1, triclinic=True)
    bt.set_stream(torch.cuda.current_stream().cuda_stream)
    for sysm in (bt, rt):
        sysm.group_create_from_indices("G", np.arange(N_ATOMS, dtype=np.uint32))
    rt.set_frames(bt.synth_blob_ref(SEED, BLOB_SCALE, [BOX / 2] * 3), tri)
    rot, cen = frame_params(0, F)
    bt.synth_blob(SEED, 0, F, BLOB_SCALE, NOISE_SCALE, rot, cen, tri, wrap=False)
    bt.atoms_wrap()  # into the triclinic cell
    t = time_op(lambda: bt.group_get_center("G", out=d_c), reps=3)
    out["triclinic_group_get_center"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "fallback_frames": bt.fallback_frames(),
                                         "frac_of_hbm_peak": 12 * N_ATOMS * F * 1e-9 / (t * 1e-3) / peak, "note": "parity unpinned (extension)"}
    t = time_op(lambda: bt.calc_rmsd(rt, "G", out=d_r), reps=3)
    rr = d_r.cpu().numpy()
    assert np.all(np.abs(rr - 0.0866) < 2e-3), rr
    out["triclinic_calc_rmsd"] = {"ms": t, "frames_per_s": F / (t * 1e-3), "fallback_frames": bt.fallback_frames(),
                                  "frac_of_hbm_peak": (12 * F + 16) * N_ATOMS * 1e-9 / (t * 1e-3) / peak, "note": "parity unpinned (extension)"}
    bt.close()
    rt.close()
    # calc_rmsd_and_fit at the headline's batch size (configs[4] names "RMSD/Kabsch fit"): 37 frames, all 4M atoms rewritten
    F37 = 37
    b = g.System(N_ATOMS, masses=m, device=local, max_frames=F37)
    r = g.System(N_ATOMS, masses=m, device=local, max_frames=1)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    for sysm in (b, r):
        sysm.group_create_from_indices("G", np.arange(N_ATOMS, dtype=np.uint32))
    r.set_frames(b.synth_blob_ref(SEED, BLOB_SCALE, [BOX / 2] * 3), [BOX] * 3)
    rot, cen = frame_params(0, F37)
    b.synth_blob(SEED, 0, F37, BLOB_SCALE, NOISE_SCALE, rot, cen, [BOX] * 3, wrap=True)
    d_r37 = torch.empty((F37,), dtype=torch.float32, device=dev)
    t = time_op(lambda: b.calc_rmsd_and_fit(r, "G", out=d_r37), reps=3)
    out["calc_rmsd_and_fit_37_frames"] = {"ms": t, "frames_per_s": F37 / (t * 1e-3),
                                          "gbs": ((12 + 24) * F37 + 16) * N_ATOMS * 1e-9 / (t * 1e-3),
                                          "frac_of_hbm_peak": ((12 + 24) * F37 + 16) * N_ATOMS * 1e-9 / (t * 1e-3) / peak}
    b.close()
    r.close()
    # configs[3]: 1M atoms, box 21.5, 2 000 x 200 000 all-pairs
    n1, n2, N = 2000, 200000, 1_000_000
    p = g.System(N, device=local, max_frames=2)
    p.set_stream(torch.cuda.current_stream().cuda_stream)
    p.group_create_from_indices("A", np.arange(n1))
    p.group_create_from_indices("B", np.arange(500000, 500000 + n2))
    p.synth_uniform(SEED, 0, 2, [-0.1 * 21.5] * 3, [1.2 * 21.5] * 3, [21.5] * 3)
    # 0.6 ms per call: enough launches for the clocks to settle after the allocations above (3 launches read 0.75 ms, 20 read 0.60 ms)
    red = {"min": torch.empty(2, dtype=torch.float32, device=dev), "argmin": torch.empty((2, 2), dtype=torch.int32, device=dev),
           "max": torch.empty(2, dtype=torch.float32, device=dev), "argmax": torch.empty((2, 2), dtype=torch.int32, device=dev),
           "count": torch.empty(2, dtype=torch.int64, device=dev)}  # results stay on the device: the kernel, not five small D2H copies
    t = time_op(lambda: p.group_all_distances_reduce("A", "B", g.Dimension.XYZ, cutoff=1.0, out=red), reps=20, warm=5)
    assert int(red["count"][0].item()) > 0
    pairs = 2 * n1 * n2
    # FP32 roofline of SURVEY 8d: 18 reference FP32 ops per pair against 148 SM x 128 lanes x 1.965 GHz = 37.2 T op/s (non-FMA)
    out["all_pairs_fused_reduce"] = {"ms": t, "pairs_per_s": pairs / (t * 1e-3), "frames_per_s": 2 / (t * 1e-3),
                                     "ref_fp32_ops_per_s": 18 * pairs / (t * 1e-3),
                                     "frac_of_fp32_peak": 18 * pairs / (t * 1e-3) / (148 * 128 * 1.965e9)}
    p.synth_uniform(SEED, 0, 1, [-0.1 * 21.5] * 3, [1.2 * 21.5] * 3, [21.5] * 3)
    mat = torch.empty((1, n1, n2), dtype=torch.float32, device=dev)
    t = time_op(lambda: p.group_all_distances("A", "B", g.Dimension.XYZ, out=mat), reps=10, warm=3)
    out["all_pairs_materialise"] = {"ms": t, "pairs_per_s": n1 * n2 / (t * 1e-3), "write_gbs": 4 * n1 * n2 * 1e-9 / (t * 1e-3),
                                    "frac_of_hbm_peak": 4 * n1 * n2 * 1e-9 / (t * 1e-3) / peak}
    del mat
    # triclinic minimum-image distances (sequential reduction + search over the 27 neighbouring images), fused reduction,
    # 2 000 x 200 000 atoms in the triclinic variant of the configs[3] box
    pt = g.System(N, device=local, max_frames=1, triclinic=True)
    pt.set_stream(torch.cuda.current_stream().cuda_stream)
    pt.group_create_from_indices("A", np.arange(n1))
    pt.group_create_from_indices("B", np.arange(500000, 500000 + n2))
    pt.synth_uniform(SEED, 0, 1, [-0.1 * 21.5] * 3, [1.2 * 21.5] * 3, [21.5, 0, 0, 5.4, 21.5, 0, 5.4, -5.4, 21.5])
    red1 = {k: v[:1] for k, v in red.items()}
    t = time_op(lambda: pt.group_all_distances_reduce("A", "B", g.Dimension.XYZ, cutoff=1.0, out=red1), reps=3, warm=1)
    out["all_pairs_triclinic_fused_reduce"] = {"ms": t, "pairs_per_s": n1 * n2 / (t * 1e-3),
                                               "note": "27-image search per pair (dodecahedron-safe), reference-order arithmetic"}
    pt.group_create_from_indices("B2", np.arange(500000, 500000 + 20000))
    mat = torch.empty((1, n1, 20000), dtype=torch.float32, device=dev)
    t = time_op(lambda: pt.group_all_distances("A", "B2", g.Dimension.XYZ, out=mat), reps=3, warm=1)
    out["all_pairs_triclinic_materialise"] = {"ms": t, "pairs_per_s": n1 * 20000 / (t * 1e-3)}
    del mat
    pt.close()
    # cutoff pair search through a cell grid: 200 000 atoms against all 1M atoms, cutoff 1.0 nm
    # (~420 neighbours per atom at 100 atoms/nm^3); the brute-force equivalent is 2e11 pairs per frame
    p.group_create_from_indices("Q", np.arange(500000, 500000 + 200000))
    p.group_create_from_indices("all1M", np.arange(N))
    cnt = [None]

    def search():
        cnt[0] = p.group_pairs_within("Q", "all1M", 1.0)[0]

    t = time_op(search, reps=3)
    found = int(cnt[0][0])
    out["pairs_within_cell_grid"] = {"ms": t, "frames_per_s": 1 / (t * 1e-3), "pairs_found_per_frame": found,
                                     "pairs_found_per_s": found / (t * 1e-3),
                                     "brute_force_equivalent_pairs_per_s": 200000 * N / (t * 1e-3)}
    # ... the same search with the pairs written out (device-resident lists, 8 B per pair; hits staged per warp in shared memory)
    d_pairs = torch.empty((1, 120_000_000, 2), dtype=torch.int32, device=dev)

    def search_store():
        cnt[0] = p.group_pairs_within("Q", "all1M", 1.0, pairs_out=d_pairs)[0]

    t = time_op(search_store, reps=3)
    out["pairs_within_cell_grid_store"] = {"ms": t, "pairs_stored_per_s": int(cnt[0][0]) / (t * 1e-3),
                                           "write_gbs": 8 * int(cnt[0][0]) * 1e-9 / (t * 1e-3)}
    del d_pairs
    p.close()
    # the two users of the cell grid on a water box of the same size and density: 333 333 rigid three-site
    # waters at random positions and orientations (1M atoms, 100 atoms / nm^3); guess_bonds with O 0.152 / H 0.12 nm radii,
    # water-water hydrogen bonds with the reference test's criteria (0.3 nm, 150 degrees; hbonds.rs:505-587)
    rng = np.random.default_rng(SEED)
    n_mol, Lw = 333_333, 21.5
    o = rng.uniform(0, Lw, size=(n_mol, 3))
    q = rng.normal(size=(n_mol, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    qw, qx, qy, qz = q.T
    R = np.stack([1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw), 2 * (qx * qy + qz * qw),
                  1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw), 2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw),
                  1 - 2 * (qx * qx + qy * qy)], axis=1).reshape(n_mol, 3, 3)
    w_xyz = np.empty((n_mol, 3, 3))
    w_xyz[:, 0] = o
    w_xyz[:, 1] = o + R @ np.array([0.0957, 0.0, 0.0])
    w_xyz[:, 2] = o + R @ np.array([-0.024, 0.0927, 0.0])
    w_xyz = np.mod(w_xyz.reshape(-1, 3), Lw).astype(np.float32)
    nw = w_xyz.shape[0]
    ws = g.System(nw, device=local, max_frames=1)
    ws.set_stream(torch.cuda.current_stream().cuda_stream)
    ws.set_frames(w_xyz, [Lw] * 3)
    vdw = np.tile(np.array([0.152, 0.12, 0.12], np.float32), n_mol)
    import ctypes as C
    from groan_rs_b200.system import _ptr
    # through the C ABI with device-resident outputs (CUDA events): the search itself, without the Python mirror's sorting
    cap_b, cap_h = 4 * nw, 2 * nw
    d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    d_bp = torch.empty((1, cap_b, 2), dtype=torch.int32, device=dev)

    def gb():
        ws._check(ws._lib.groan_gpu_guess_bonds(ws._h, _ptr(vdw), C.c_float(0.55), _ptr(d_cnt), _ptr(d_bp), cap_b), "guess_bonds")

    t = time_op(gb, reps=3)
    out["guess_bonds_water_1M"] = {"ms": t, "frames_per_s": 1 / (t * 1e-3), "bonds": int(d_cnt.item()),
                                   "note": "groan_gpu_guess_bonds, 1M atoms: 4 MB of radii uploaded per call, grid build, search, pairs left on the device"}
    ow = np.arange(0, nw, 3)
    ws.group_create_from_indices("OW", ow)
    don = ow.astype(np.uint32)
    off = (2 * np.arange(n_mol + 1)).astype(np.uint32)
    hyd = np.stack([ow + 1, ow + 2], axis=1).reshape(-1).astype(np.uint32)
    d_dha = torch.empty((1, cap_h, 3), dtype=torch.int32, device=dev)
    d_da = torch.empty((1, cap_h, 2), dtype=torch.float32, device=dev)

    def hb():
        ws._check(ws._lib.groan_gpu_hbonds(ws._h, ws._gid("OW"), _ptr(don), _ptr(off), _ptr(hyd), int(don.size), C.c_float(0.3), C.c_float(150.0),
                                           _ptr(d_cnt), _ptr(d_dha), _ptr(d_da), cap_h), "hbonds")

    t = time_op(hb, reps=3)
    out["hbonds_water_1M"] = {"ms": t, "frames_per_s": 1 / (t * 1e-3), "hbonds": int(d_cnt.item()),
                              "note": "groan_gpu_hbonds, 333 333 donors with two hydrogens each against 333 333 acceptors, 0.3 nm / 150 degrees "
                                      "(random waters: few bonds qualify); donor lists uploaded per call, records left on the device"}
    ws.close()
    return out


def main():
    # stdout carries exactly one JSON line: route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=37,
                    help="frames per step (48 MB each); 37 frames x 8 CTAs per frame = 296 CTAs = one wave of 2 CTAs per SM")
    ap.add_argument("--cpu-frames", type=int, default=None, help="distinct frames held by the CPU arm (default: frames per step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
