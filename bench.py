#!/usr/bin/env python
"""bench.py -- tracked frames/s (5 templates/frame) of the Laser Spot Track hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config B]

One process per GPU (the driver launches N>1 through torch.distributed.run).  A *step* is one
pass of the hot path over one batch of ``--frames`` synthetic slices of the named configuration
(default: BASELINE.json configs[1] -- 1920x1080 16-bit, 48x48 templates, 160x160 search ROIs).
Frames are sharded over the ranks (each rank tracks its own contiguous chunk; nothing is reduced
between GPUs, so there is no data-path collective); ``scaling`` is "weak".

Two legs are timed, each with W warm-up steps and exactly K timed steps bracketed by a barrier +
torch.cuda.synchronize() on both sides, CUDA events on the device, max over ranks:
  value : slices already resident in HBM (a ring of distinct slices larger than L2)
  e2e   : the public API with HOST buffers -- pinned slices, host->device copies and the
          device->host read of the per-frame records inside the timed region.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from laser_spot_track_b200.synth import CONFIGS, StackConfig, SyntheticStack  # noqa: E402

METRIC = "tracked frames/sec (5 templates/frame)"
UNIT = "frames/s"
WORKLOADS = {
    "A": "synthetic 640x480 8-bit stack, 4 marks + spot, 32x32 templates, 96x96 search ROIs",
    "B": "1920x1080 16-bit timelapse, 48x48 templates, 160x160 search ROIs",
    "C": "4096x3072 16-bit stack, whole-slice search (sArea = 0) of all five 64x64 templates",
    "D": "large-template stress: 2048x2048 8-bit, 128x128 marks and spot over 512x512 ROIs",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="B", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="slices per step per GPU (0 = per-config default)")
    ap.add_argument("--ring", type=int, default=0, help="distinct slices in the ring (0 = per-config default)")
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="frames of the CPU baseline sample (0 = default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short device-resident runs of configs A, C and D")
    ap.add_argument("--lanes", type=int, default=0,
                    help="tracking lanes per GPU (lst_multi_create with the device listed this many times): the host-side "
                         "chain walk of one lane overlaps the kernels of the others; 0 = up to %d, fewer when the host has "
                         "less than one core per lane thread" % LANES_DEFAULT)
    ap.add_argument("--continuous", action="store_true",
                    help="device leg: the steps are consecutive pieces of ONE endless stack (slice numbers run on across ranks and "
                         "steps), so with --frames not a multiple of the ring the chunk boundaries fall at different phases of the "
                         "motion, the speculated boundary states are often wrong and chunks are re-tracked (resolver.chunks_retracked)")
    ap.add_argument("--lead-in", type=int, default=0,
                    help="--continuous: slices of the previous chunk a rank tracks first to refine its speculated start state")
    ap.add_argument("--no-files", action="store_true", help="skip the slices-from-TIFF-files leg (e2e_files)")
    ap.add_argument("--ingest", type=int, default=0, choices=[0, 1, 2], help="0 ROI DMA (default), 1 whole slices, 2 ROI gather kernel")
    ap.add_argument("--engine", default="auto", choices=["auto", "dp4a", "tc"], help="match kernel: tensor cores (default when the template fits) or IDP.4A")
    return ap.parse_args()


def config_dict(name: str, cfg) -> dict:
    """The workload as BOTH arms name it (identical keys and values: the driver compares them)."""
    return {"workload": WORKLOADS[name], "frame": [cfg.width, cfg.height], "depth": cfg.dtype,
            "templ_size": cfg.templ_size, "s_area": cfg.s_area}


def bench_config(name: str, ring: int) -> StackConfig:
    c = CONFIGS[name]
    return StackConfig(c.width, c.height, c.dtype, c.templ_size, c.s_area, n_frames=ring, motion="periodic", period=ring,
                       blob_variety=c.blob_variety)


# DRAM bytes per window of the match kernel from the committed ncu --set full captures (profiles/):
# lst_ncc_match (IDP.4A), config B: 66.9 MB for 2560 windows of 160x160 8-bit pixels (the window is
#   read once; 25.6 KB algorithmic) -- profiles/r1_ncu_full_lst_ncc_match_configB.txt;
# lst_ncc_tc (tensor core), config B: 121.6 MB read + 4.3 MB written for 4704 windows = 26.8 KB per window
#   (1.05 x the 25.6 KB of the 8-bit window: the window sums never leave the SM any more) --
#   profiles/r2_ncu_full_three_kernels_configB.txt.
NCU_DRAM_BYTES_PER_WINDOW = {("B", False): 66.9e6 / 2560, ("B", True): (121.614848e6 + 4.339712e6) / 4704}

LANES_DEFAULT = 4

DEFAULTS = {  # frames per step, ring length, max_batch, cpu sample
    "A": (8192, 256, 2048, 512),
    "B": (10240, 64, 1024, 128),
    "C": (16, 8, 4, 2),
    "D": (64, 32, 64, 8),
}


# ---------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's OpenCV-backed loop on host cores
# ---------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    cfg_name, ring, lo, hi = args
    from oracle import lst_oracle as O
    import cv2
    cv2.setNumThreads(1)
    cfg = bench_config(cfg_name, ring)
    st = SyntheticStack(cfg)
    op = O.Params(cfg.width, cfg.height, cfg.templ_size, cfg.s_area)
    tr = O.OracleTracker(op, backend="cv2")
    tr.set_reference(st.frame(0), st.ref_xy)
    frames = [st.frame(i) for i in range(lo, hi)]       # generation excluded from the timing
    t0 = time.perf_counter()
    recs = tr.run(frames)
    dt = time.perf_counter() - t0
    return hi - lo, dt, sum(r.status == 0 for r in recs)


def cpu_reference_rate(cfg_name: str, ring: int, sample: int, procs: int):
    """frames/s of the CPU path on `procs` host cores over `sample` frames.  Chunks start at
    multiples of the ring period, where the periodic scene is back at its reference pose, so every
    chunk legitimately starts from the reference displacement state."""
    if procs <= 1:
        n, dt, ok = _cpu_chunk((cfg_name, ring, 1, 1 + sample))
        return n / dt, n, ok
    import multiprocessing as mp
    per = max(sample // procs, 16)
    jobs = [(cfg_name, ring, 1 + (i * ring) % max(ring, 1), 1 + (i * ring) % max(ring, 1) + per) for i in range(procs)]
    with mp.get_context("fork").Pool(procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_chunk, jobs)
        wall = time.perf_counter() - t0
    n = sum(r[0] for r in res)
    busy = max(r[1] for r in res)            # chunks run concurrently: the slowest one bounds the job
    return n / busy, n, sum(r[2] for r in res)


def run_reference(args, rank, world):
    if rank != 0:
        return
    frames_d, ring_d, _, sample_d = DEFAULTS[args.config]
    ring = args.ring or ring_d
    procs = os.cpu_count() or 1
    sample = args.cpu_sample or sample_d
    per_step = max(sample, procs * 16)
    vals = []
    for i in range(args.warmup + args.steps):
        rate, n, ok = cpu_reference_rate(args.config, ring, per_step, procs)
        if i >= args.warmup:
            vals.append((rate, n))
    rate = float(np.mean([v[0] for v in vals]))
    nfr = vals[0][1]
    cfg = bench_config(args.config, ring)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * nfr / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args.config, cfg), "details": {"frames_per_step": nfr},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{nfr} frames per step on {procs} processes: oracle loop (numpy ImageJ blur + "
                                   f"cv2.matchTemplate TM_CCOEFF_NORMED, IPP off, 1 thread each); the Java plugin "
                                   f"itself cannot run here (no JVM)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from laser_spot_track_b200 import LaserSpotTrack4, LaserSpotTrack4Multi, native as N

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the tracking path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # Host placement: every rank gets its own slice of the cores this process may use, so that the lane threads of one
    # rank (they poll their streams) never preempt another rank's.  LST_BENCH_NO_PIN=1 leaves the scheduler alone.
    host_cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else list(range(os.cpu_count() or 1))
    my_cores = host_cores
    if world > 1 and not os.environ.get("LST_BENCH_NO_PIN") and len(host_cores) >= 2 * world:
        per = len(host_cores) // world
        my_cores = host_cores[local_rank * per:(local_rank + 1) * per]
        try:
            os.sched_setaffinity(0, my_cores)
        except OSError:
            my_cores = host_cores
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")     # alternative carrier of the chunk-boundary hand-off (LST_BENCH_EXCHANGE=gloo)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frames_d, ring_d, mb_d, sample_d = DEFAULTS[args.config]
    F = args.frames or frames_d
    ring = args.ring or ring_d
    cfg = bench_config(args.config, ring)
    st = SyntheticStack(cfg)
    bits = {"u8": 8, "u16": 16, "f32": 32}[cfg.dtype]
    fb = cfg.width * cfg.height * (bits // 8)

    lib = N.load()
    trk = LaserSpotTrack4(cfg.width, cfg.height, bits, device=local_rank, templSize=cfg.templ_size, sArea=cfg.s_area,
                          max_batch=args.max_batch or mb_d, ncc_engine={"auto": 0, "dp4a": 1, "tc": 2}[args.engine],
                          ingest_mode=args.ingest)
    env_engine = os.environ.get("LST_NCC_ENGINE")
    use_tc = (args.engine != "dp4a" and env_engine != "dp4a" and cfg.templ_size <= 181) or env_engine == "tc"
    # ring of distinct slices: pinned host copy (e2e leg) and HBM copy (value leg)
    host_ring = C.c_void_p()
    assert lib.lst_alloc_pinned(ring * fb, C.byref(host_ring)) == 0
    hr = np.ctypeslib.as_array(C.cast(host_ring, C.POINTER(C.c_uint8)), shape=(ring * fb,))
    for i in range(ring):
        hr[i * fb:(i + 1) * fb] = st.frame(i).view(np.uint8).reshape(-1)
    dev_ring = trk.device_alloc(ring * fb)
    trk._check(lib.lst_device_upload(trk._ctx, C.c_void_p(dev_ring), host_ring, ring * fb))
    info = trk.set_reference(st.frame(0), st.ref_xy)
    # every lane is a host thread that waits on its stream: keep one host core per lane across the ranks of this node
    # one core per lane thread (this rank's main thread sleeps while the lanes run)
    lanes = args.lanes if args.lanes > 0 else max(1, min(LANES_DEFAULT, len(my_cores)))
    if args.lanes <= 0 and F // (args.max_batch or mb_d) < 4:
        lanes = 1       # a lane should get several full batches; short steps (config D: 64 slices) stay on one
    if args.lanes <= 0 and cfg.s_area == 0:
        lanes = 1       # whole-slice search: no recurrence to walk and every pass fills the GPU (4 lanes: 1.31 k, 1 lane: 1.70 k slices/s)
    runner = trk
    if lanes > 1:
        runner = LaserSpotTrack4Multi(cfg.width, cfg.height, bits, [local_rank] * lanes, templSize=cfg.templ_size,
                                      sArea=cfg.s_area, max_batch=args.max_batch or mb_d,
                                      ncc_engine={"auto": 0, "dp4a": 1, "tc": 2}[args.engine], ingest_mode=args.ingest)
        runner.set_reference(st.frame(0), st.ref_xy)

    # this rank's chunk of each step: F consecutive slices of the (cyclic) stack, starting at slice 1
    idx = [(1 + j) % ring for j in range(F)]
    dev_ptrs = (C.c_void_p * F)(*[dev_ring + i * fb for i in idx])
    host_ptrs = (C.c_void_p * F)(*[host_ring.value + i * fb for i in idx])
    out = (N.Record * (F + max(0, args.lead_in)))()      # --continuous --lead-in: the lead-in slices' records come first

    # Sharding: the step's stack is world*F slices; rank r owns the contiguous chunk r.  The state
    # at the chunk boundary is speculated (the ring is periodic, so "my own previous end state" is
    # the natural guess), exchanged host-side and verified (laser_spot_track_b200/sharded.py).
    from laser_spot_track_b200.sharded import ShardedTracker, TorchComm

    class _Adapter:
        def __init__(self, fn):
            self.fn = fn

        def set_state(self, d):
            runner.set_state(d)

        def get_state(self):
            return runner.get_state()

        def track(self, frames, first):
            return self.fn(frames, first, out)

    # Boundary hand-off between the ranks: 21 doubles per rank and step -- control information, not slice data; nothing
    # on the data path is reduced between GPUs.  They travel as a tiny device tensor through the NCCL group the timing
    # barrier uses anyway: measured on 8 GPUs 6.46 M frames/s against 6.21 M with a gloo (TCP loopback) group, whose
    # helper threads compete with the lane threads for this rank's host cores.  LST_BENCH_EXCHANGE=gloo switches back.
    if world > 1 and os.environ.get("LST_BENCH_EXCHANGE", "nccl") == "nccl":
        comm = TorchComm(None, torch.device("cuda", local_rank))
    else:
        comm = TorchComm(host_group, "cpu") if world > 1 else None
    if os.environ.get("LST_BENCH_NO_COMM"):      # diagnosis only: no boundary exchange (each rank trusts its own guess)
        comm = None
    sh_dev = ShardedTracker(_Adapter(runner.track_device_ptrs), comm)
    sh_e2e = ShardedTracker(_Adapter(runner.track_host_ptrs), comm)

    # N > 1: the exchange of a step's boundary states (21 doubles, gloo) travels while the next step -- an independent
    # stack -- is already being tracked (ShardedTracker.track_chunk_pipelined); the records of a step are final when its
    # exchange has completed, which for the last step happens inside the timed region (finish()).  The two record
    # buffers alternate so that a late re-track would find its records intact.
    out_alt = (N.Record * (F + max(0, args.lead_in)))() if world > 1 else out
    flip = [0]

    def _next_out():
        flip[0] ^= 1
        return out if flip[0] else out_alt

    class _AdapterAlt(_Adapter):
        def track(self, frames, first):
            return self.fn(frames, first, _next_out())

    pipelined = world > 1 and comm is not None and bool(os.environ.get("LST_BENCH_PIPELINED_EXCHANGE"))
    if pipelined:
        sh_dev = ShardedTracker(_AdapterAlt(runner.track_device_ptrs), comm)
        sh_e2e = ShardedTracker(_AdapterAlt(runner.track_host_ptrs), comm)

    cont = {"t": 0, "ptrs": []}
    LEAD = max(0, args.lead_in)
    if args.continuous:      # pointer arrays of the steps of the device leg, built before the timed region
        for t in range(args.steps + max(args.warmup, 0) + 2):
            g0 = 1 + (t * world + rank) * F
            # with a lead-in the array starts with the last LEAD slices of the previous chunk: tracked in the same call from the
            # speculated state, they pull it onto the true one (ShardedTracker.track_chunk(..., lead_in_count=LEAD))
            cont["ptrs"].append((C.c_void_p * (F + LEAD))(*[dev_ring + ((g0 - LEAD + j) % ring) * fb for j in range(F + LEAD)]))

    def step_device():
        ptrs, first, nlead = dev_ptrs, 1 + rank * F, 0
        if args.continuous:
            t = cont["t"]
            cont["t"] = t + 1
            ptrs, first, nlead = cont["ptrs"][t % len(cont["ptrs"])], 1 + (t * world + rank) * F, LEAD
        if pipelined:
            sh_dev.track_chunk_pipelined(ptrs, first, runner.get_state(), None, nlead)
        else:
            sh_dev.track_chunk(ptrs, first, runner.get_state(), None, nlead)

    def step_e2e():
        if pipelined:
            sh_e2e.track_chunk_pipelined(host_ptrs, 1 + rank * F, runner.get_state())
        else:
            sh_e2e.track_chunk(host_ptrs, 1 + rank * F, runner.get_state())

    def _finish_exchanges():
        for sh in (sh_dev, sh_e2e):
            if getattr(sh, "_pending", None) is not None:
                sh.finish()

    def timed(step_fn, K, W, who=None):
        who = who or runner
        for _ in range(W):
            step_fn()
        _finish_exchanges()
        who.reset_stats()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            step_fn()
        _finish_exchanges()       # the last step's boundary exchange completes inside the timed region
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = who.stats()
        return float(t[0]), float(t[1]), s

    K, W = args.steps, max(args.warmup, 0)
    sampler = ClockSampler(local_rank)
    peaks = (C.c_double * 3)()
    have_peaks = lib.lst_measure_alu_peaks(local_rank, 20.0, peaks) == 0

    if lanes == 1:
        trk.set_kernel_timing(True)
    if rank == 0:
        sampler.start()
    ms_dev, wall_dev, s_dev = timed(step_device, K, W)
    # chunks tracked a second time because their speculated start state placed a window differently (warm-up steps
    # included): between ranks and between the lanes of a rank, summed over the ranks
    retr_t = torch.tensor([float(sh_dev.retracked), float(runner.retracked) if lanes > 1 else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(retr_t, op=dist.ReduceOp.SUM)
    retr_all = [int(v) for v in retr_t.tolist()]
    clocks = sampler.stop() if rank == 0 else None
    ok_frames = sum(1 for i in range(F) if out[i].status == 0)      # rank 0 reports: its records start the buffer
    trk.set_kernel_timing(False)
    # Per-launch kernel times need kernels that do not overlap: with several lanes the roofline numbers come
    # from a second timed region of the same K steps on one lane (same kernels, same launches).
    s_roof, ms_roof = s_dev, ms_dev
    if lanes > 1:
        def step_single():
            trk.track_device_ptrs(dev_ptrs, 1 + rank * F, out)
        trk.set_kernel_timing(True)
        ms_roof, _, s_roof = timed(step_single, K, 1, trk)
        trk.set_kernel_timing(False)

    def pcie_peak_gbs():
        """Measured host->device copy bandwidth of this rank (256 MB pinned cudaMemcpyAsync, all ranks
        at the same time): the denominator of the end-to-end leg, which is ingest bound."""
        nbytes = 256 << 20
        src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return 4 * nbytes / (float(t[0]) * 1e-3) / 1e9

    e2e = None
    if not args.no_e2e:
        ms_e2e, wall_e2e, s_e2e = timed(step_e2e, K, W)
        pcie = pcie_peak_gbs()
        e2e = {"value": world * F * K / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(s_e2e.h2d_bytes // K), "d2h_bytes_per_step": int(s_e2e.d2h_bytes // K),
               "ms_per_step": ms_e2e / K, "roi_fallback_windows_per_step": int(s_e2e.roi_fallback_tasks // K),
               "host_ms_per_step": s_e2e.host_ms / K, "device_wait_ms_per_step": s_e2e.device_wait_ms / K,
               "passes_per_step": s_e2e.passes / K, "tasks_per_step": s_e2e.tasks / K,
               "pcie": {"achieved_gbs_per_gpu": s_e2e.h2d_bytes / K / (ms_e2e / K * 1e-3) / 1e9,
                        "peak_gbs_per_gpu": pcie, "frac": s_e2e.h2d_bytes / K / (ms_e2e / K * 1e-3) / 1e9 / pcie,
                        "how": "peak = 256 MB pinned cudaMemcpyAsync H2D, every rank at the same time, slowest rank",
                        "ceiling_fps": world * pcie * 1e9 / max(1.0, s_e2e.h2d_bytes / K / F),
                        "ceiling": "what this box's host side can deliver: all ranks' concurrent pinned-copy rate divided by the bytes "
                                   "a slice sends (search regions + margin); the leg cannot scale past it whatever the GPUs do"},
               "ingest": "pinned host ring of whole slices; only the five search regions (+8 px margin, 32-byte aligned) of "
                         "each slice cross PCIe, as strided 3-D DMA copies on a copy stream while the previous batch computes"}

    # The same stack as files: one big-endian 16-bit TIFF per slice, as ImageJ writes them,
    # in a temporary directory (page cache resident); lst_track_files maps each file and reads only the
    # rows of the search regions.  Reported beside e2e; N = 1 only.
    e2e_files = None
    if not args.no_e2e and not args.no_files and world == 1:
        import shutil
        import struct
        import tempfile
        need = ring * (fb + 4096)
        shm_ok = os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * need
        tdir = tempfile.mkdtemp(prefix="lst_bench_", dir="/dev/shm" if shm_ok else None)
        try:
            def write_tiff(path, img):
                h, w = img.shape
                be = img.dtype.itemsize > 1
                bo = ">" if be else "<"
                raw = img.astype(img.dtype.newbyteorder(bo)).tobytes()
                ent = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, img.dtype.itemsize * 8), (259, 3, 1, 1), (262, 3, 1, 1),
                       (273, 4, 1, 8), (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(raw)),
                       (339, 3, 1, 3 if img.dtype.kind == "f" else 1)]
                ifd = struct.pack(bo + "H", len(ent))
                for tag, typ, cnt, val in ent:
                    ifd += struct.pack(bo + "HHI", tag, typ, cnt) + (struct.pack(bo + "HH", val, 0) if typ == 3 else struct.pack(bo + "I", val))
                ifd += struct.pack(bo + "I", 0)
                with open(path, "wb") as f:
                    f.write((b"MM" if be else b"II") + struct.pack(bo + "HI", 42, 8 + len(raw)) + raw + ifd)
            files = []
            for i in range(ring):
                fp = os.path.join(tdir, f"slice{i:05d}.tif")
                write_tiff(fp, st.frame(i))
                files.append(fp)
            paths = [files[i] for i in idx]

            def step_files():
                trk.track_files(paths, first_frame_index=1)

            ms_f, wall_f, s_f = timed(step_files, K, 1, trk)
            e2e_files = {"value": F * K / (wall_f * 1e-3), "unit": UNIT, "ms_per_step": wall_f / K,
                         "h2d_bytes_per_step": int(s_f.h2d_bytes // K), "file_bytes": int(os.path.getsize(files[0])),
                         "reader_threads": int(os.environ.get("LST_FILE_THREADS", min(16, os.cpu_count() or 1))), "host_cores": os.cpu_count(),
                         "what": "lst_track_files over %d paths per step cycling %d big-endian TIFF files in %s (page cache); "
                                 "wall clock, includes mapping and parsing every file" % (F, ring, os.path.dirname(files[0]))}
        except (OSError, RuntimeError) as e:      # e.g. no room for the slice files: the leg is optional
            e2e_files = {"value": None, "unit": UNIT, "error": str(e)[:200]}
        finally:
            shutil.rmtree(tdir, ignore_errors=True)

    # The same scene as a camera time-lapse of JPEG files (the EXIF path LST4:1199-1223): 4:2:0 baseline JPEGs
    # of the 8-bit scene (RGB slices, matched on intensity).  Entropy decoding is sequential per file and spreads over the
    # reader threads; only the MCUs under the search regions are reconstructed.  Bounded: 1024 slices, one step.
    e2e_jpeg = None
    if not args.no_e2e and not args.no_files and world == 1 and args.config == "B":
        import shutil
        import tempfile
        try:
            import cv2
            jdir = tempfile.mkdtemp(prefix="lst_bench_jpg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            try:
                jfiles = []
                nj = min(ring, 32)
                for i in range(nj):
                    g = (st.frame(i) >> 8).astype(np.uint8) if st.frame(i).dtype == np.uint16 else st.frame(i).astype(np.uint8)
                    fp = os.path.join(jdir, f"shot{i:05d}.jpg")
                    cv2.imwrite(fp, np.stack([g, g, g], axis=-1), [cv2.IMWRITE_JPEG_QUALITY, 92, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                                   cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420])
                    jfiles.append(fp)
                FJ = min(F, 1024)
                jpaths = [jfiles[i % nj] for i in range(1, FJ + 1)]
                with LaserSpotTrack4(cfg.width, cfg.height, 24, device=local_rank, templSize=cfg.templ_size, sArea=cfg.s_area,
                                     max_batch=args.max_batch) as jt:
                    jt.set_reference_file(jfiles[0], st.ref_xy)
                    jt.track_files(jpaths)          # warm-up at full size: the pinned staging is allocated here, not in the timed run
                    jt.set_state([0.0] * 10)
                    t0 = time.perf_counter()
                    recs = jt.track_files(jpaths)
                    wall_j = (time.perf_counter() - t0) * 1e3
                    e2e_jpeg = {"value": FJ / (wall_j * 1e-3), "unit": UNIT, "ms_per_step": wall_j, "frames": FJ,
                                "accepted": int(sum(1 for r in recs if r.status == 0)),
                                "file_bytes": int(os.path.getsize(jfiles[1])),
                                "reader_threads": int(os.environ.get("LST_FILE_THREADS", min(16, os.cpu_count() or 1))),
                                "what": "lst_track_files over %d paths cycling %d 4:2:0 JPEG files of the same scene (RGB slices, matchIntensity); "
                                        "wall clock; Huffman decoding of every file on the reader threads" % (FJ, nj)}
            finally:
                shutil.rmtree(jdir, ignore_errors=True)
        except Exception as e:      # the leg is optional
            e2e_jpeg = {"value": None, "unit": UNIT, "error": str(e)[:200]}

    # The same stack in PAGEABLE host memory (what a JNI shim holding Java arrays hands over): the copy engine
    # cannot gather from it, so reader threads pack the rows of the search regions into pinned staging.
    e2e_pageable = None
    if not args.no_e2e and world == 1:
        page_ring = np.empty((ring, fb), dtype=np.uint8)
        for i in range(ring):
            page_ring[i] = st.frame(i).view(np.uint8).reshape(-1)
        page_ptrs = (C.c_void_p * F)(*[page_ring[i].ctypes.data for i in idx])

        def step_pageable():
            sh_pg.track_chunk(page_ptrs, 1 + rank * F, runner.get_state())

        sh_pg = ShardedTracker(_Adapter(runner.track_host_ptrs), comm)
        ms_p, wall_p, s_p = timed(step_pageable, max(1, K // 2), 1)
        kp = max(1, K // 2)
        e2e_pageable = {"value": F * kp / (wall_p * 1e-3), "unit": UNIT, "ms_per_step": wall_p / kp,
                        "h2d_bytes_per_step": int(s_p.h2d_bytes // kp),
                        "what": "slices in pageable host memory; rows of the five search regions packed into pinned staging "
                                "by reader threads, one contiguous DMA per region and batch (ingest_mode 0 on pageable memory)"}
        del page_ptrs, page_ring

    value = world * F * K / (ms_dev * 1e-3)
    win_bytes_per_frame = (5 * (cfg.templ_size + 2 * cfg.s_area) ** 2 if cfg.s_area > 0 else cfg.width * cfg.height) * (bits // 8)
    # roofline of the dominant kernel (the template match): exact u8 x u8 -> s32 multiply-accumulates
    if cfg.s_area > 0:
        win_macs = float((2 * cfg.s_area + 1) ** 2 * cfg.templ_size ** 2)
    else:
        win_macs = float((cfg.width - cfg.templ_size + 1) * (cfg.height - cfg.templ_size + 1) * cfg.templ_size ** 2)
    macs_per_launch = s_roof.ncc_algorithmic_macs / max(s_roof.ncc_kernel_launches, 1)
    kern_ms = s_roof.ncc_kernel_ms / max(s_roof.ncc_kernel_launches, 1)
    achieved = 2.0 * macs_per_launch / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else None
    dp4a_peak = 2.0 * peaks[0] / 1e12 if have_peaks else None
    imma_peak = 2.0 * peaks[2] / 1e12 if (have_peaks and peaks[2] > 0) else None
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        measured = {}
    if use_tc:
        # tensor-core engine: exact u8 x u8 -> s32 Toeplitz GEMM on tcgen05.mma kind::i8.  Denominator: the
        # int8 tensor peak measured live (MEASURED_PEAKS.json only has the bf16 figure; int8 is nominally 2x)
        peak = imma_peak
        roofline = {
            "bound": "tensor", "kernel": "lst_ncc_tc", "achieved": achieved, "peak": peak, "unit": "TOP/s",
            "frac": (achieved / peak) if (achieved and peak) else None,
            "peak_source": "measured live: tcgen05.mma.kind::i8 M128 N256 K32 microbenchmark (lst_measure_alu_peaks); "
                           "MEASURED_PEAKS.json bf16_tflops=%s for comparison" % measured.get("bf16_tflops"),
            "note": "achieved counts ALGORITHMIC ops (2*R^2*T^2 per window); the Toeplitz embedding executes ~2.6x more "
                    "MACs and small-N MMAs are bound by the shared-memory read of the A operand, so the fraction of the "
                    "tensor peak is low while the kernel is several times faster than the IDP.4A engine, whose own "
                    "roofline is given as dp4a_engine_peak",
            "dp4a_engine_peak": dp4a_peak, "achieved_over_dp4a_peak": (achieved / dp4a_peak) if (achieved and dp4a_peak) else None,
        }
        # The bound that applies to this embedding: tcgen05.mma.kind::i8 reads both operands from shared memory at
        # 128 B/clk/SM (tools/tc_rate.cu: an M128 K32 MMA takes 40 cycles at N = 32 and 64 at N = 128 -- exactly
        # (4 KB + N * 32 B) / 128).  Per template row the window tile's K steps are read once (4 KB each) and the
        # Toeplitz band blocks they feed (1 KB each).
        if cfg.s_area > 0:
            R = 2 * cfg.s_area + 1
            nd = (cfg.templ_size + 62) // 32
            nbe = min(4, (R + 31) // 32)
            nk = nbe + nd - 1
            bblocks = sum(min(nbe - 1, j) - max(0, j - nd + 1) + 1 for j in range(nk))
            tiles = ((R + 127) // 128) * ((R + 32 * nbe - 1) // (32 * nbe))
            smem_bytes_per_window = tiles * cfg.templ_size * (nk * 4096 + bblocks * 1024)
            windows_per_launch = macs_per_launch / max(win_macs, 1)
            sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
            floor_ms = windows_per_launch * smem_bytes_per_window / 128.0 / (148 * sm_hz) * 1e3
            roofline["smem_operand_floor"] = {
                "bytes_per_window": smem_bytes_per_window, "floor_ms_per_launch": floor_ms,
                "frac_of_floor": floor_ms / kern_ms if kern_ms > 0 else None,
                "what": "time the tensor core needs just to read its operands from shared memory (128 B/clk/SM) for the "
                        "launch's windows; the vertical sums and the epilogue of the worker warps share that pipe"}
    else:
        peak = dp4a_peak
        roofline = {
            "bound": "int8-dp4a-pipe", "kernel": "lst_ncc_match", "achieved": achieved, "peak": peak, "unit": "TOP/s",
            "frac": (achieved / peak) if (achieved and peak) else None,
            "peak_source": "measured live: IDP.4A.U8.U8 microbenchmark (lst_measure_alu_peaks); MEASURED_PEAKS.json "
                           "holds only HBM and bf16-tensor peaks and this kernel is bound by neither (SURVEY.md 8d)",
            "int8_tensor_peak": imma_peak,
        }
    roofline.update({
        "traffic": (NCU_DRAM_BYTES_PER_WINDOW.get((args.config, use_tc)) or 0) * (macs_per_launch / max(win_macs, 1)) or None,
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel "
                          "(profiles/), per window, times the windows of an average launch",
        "ffma_peak_tflops": 2.0 * peaks[1] / 1e12 if have_peaks else None,
        "algorithmic_ops_per_launch": 2.0 * macs_per_launch, "avg_launch_ms": kern_ms,
        "kernel_share_of_step": s_roof.ncc_kernel_ms / ms_roof if ms_roof > 0 else None,
        "launches_timed": int(s_roof.ncc_kernel_launches),
        "kernel_ms_per_step": {"lst_preprocess": s_roof.preprocess_ms / K, "match": s_roof.ncc_kernel_ms / K,
                               "lst_epilogue": s_roof.epilogue_ms / K, "sum": (s_roof.preprocess_ms + s_roof.ncc_kernel_ms + s_roof.epilogue_ms + s_roof.box_sums_ms) / K,
                               "step": ms_roof / K},
        # every kernel of the step against the unit that bounds it, recomputable from the numbers beside it
        "per_kernel": {
            "lst_preprocess": {
                "bound": "fp32-issue", "unit": "TFLOP/s",
                "flop_per_px": 38, "px_per_step": s_roof.preprocess_px / K, "ms_per_step": s_roof.preprocess_ms / K,
                "achieved": (38.0 * s_roof.preprocess_px / (s_roof.preprocess_ms * 1e-3) / 1e12) if s_roof.preprocess_ms > 0 else None,
                "peak": (peaks[1] / 1e12) if have_peaks else None,
                "frac": (38.0 * s_roof.preprocess_px / (s_roof.preprocess_ms * 1e-3) / peaks[1]) if (have_peaks and s_roof.preprocess_ms > 0 and peaks[1] > 0) else None,
                "peak_source": "measured live FFMA instruction rate (lst_measure_alu_peaks): ImageJ's blur multiplies and adds "
                               "separately in Java's float order, so a flop is an instruction -- half the FMA flop peak"},
            "match": {"bound": roofline["bound"], "unit": "TOP/s", "achieved": achieved, "peak": peak,
                      "frac": roofline["frac"], "ms_per_step": s_roof.ncc_kernel_ms / K},
            "lst_epilogue": {"bound": "latency", "ms_per_step": s_roof.epilogue_ms / K,
                             "what": "one small block per window: sub-pixel fit and accept test from the nine scores the match kernel left"},
        },
        "measured_with": "the timed region of `value`" if lanes == 1 else
                         "a second timed region of the same %d steps on ONE lane (per-launch CUDA-event times need kernels that "
                         "do not overlap; `value` runs %d lanes): %.0f frames/s" % (K, lanes, world * F * K / (ms_roof * 1e-3)),
        "window_bytes_per_frame": win_bytes_per_frame,
        "hbm_view": {"peak_gbs": measured.get("hbm_gbs"), "window_bytes_per_s": value * win_bytes_per_frame / 1e9,
                     "frac": (value * win_bytes_per_frame / 1e9 / measured["hbm_gbs"]) if measured.get("hbm_gbs") else None,
                     "what": "raw window bytes the blur kernel reads per second against the HBM copy peak: the path is not HBM bound"},
    })

    # the other match engine on the same device-resident workload, for its roofline (not the headline)
    roofline_alt = None
    if world == 1 and args.engine == "auto" and env_engine is None and cfg.templ_size <= 181:
        alt = LaserSpotTrack4(cfg.width, cfg.height, bits, device=local_rank, templSize=cfg.templ_size, sArea=cfg.s_area,
                              max_batch=args.max_batch or mb_d, ncc_engine=1)
        alt.set_reference(st.frame(0), st.ref_xy)
        alt.set_kernel_timing(True)
        out2 = (N.Record * F)()
        for _ in range(2):
            alt.track_device_ptrs(dev_ptrs, 1, out2)
        alt.reset_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        alt.track_device_ptrs(dev_ptrs, 1, out2)
        dt = time.perf_counter() - t0
        sa = alt.stats()
        a_ach = 2.0 * sa.ncc_algorithmic_macs / (sa.ncc_kernel_ms * 1e-3) / 1e12 if sa.ncc_kernel_ms > 0 else None
        roofline_alt = {"bound": "int8-dp4a-pipe", "kernel": "lst_ncc_match", "achieved": a_ach, "peak": dp4a_peak, "unit": "TOP/s",
                        "frac": (a_ach / dp4a_peak) if (a_ach and dp4a_peak) else None,
                        "value_with_this_engine": F / dt, "identical_results": all(
                            (a.status, a.x_abs, a.y_abs, a.dL) == (b.status, b.x_abs, b.y_abs, b.dL) for a, b in zip(out, out2))}
        alt.close()

    # The other named configurations, device resident, a few steps each (one context, one lane): so that every BASELINE
    # config has a number from the same run.  Their full-geometry parity tests are tests/test_gpu_configs.py.
    extra = None
    if world == 1 and args.config == "B" and not args.no_extra:
        extra = {}
        for name in ("A", "C", "D"):
            try:
                xf, xring, xmb, _ = DEFAULTS[name]
                xcfg = bench_config(name, xring)
                xst = SyntheticStack(xcfg)
                xbits = {"u8": 8, "u16": 16, "f32": 32}[xcfg.dtype]
                xfb = xcfg.width * xcfg.height * (xbits // 8)
                xt = LaserSpotTrack4(xcfg.width, xcfg.height, xbits, device=local_rank, templSize=xcfg.templ_size, sArea=xcfg.s_area,
                                     max_batch=xmb)
                xdev = xt.device_alloc(xring * xfb)
                for i in range(xring):
                    fr = np.ascontiguousarray(xst.frame(i))
                    xt._check(lib.lst_device_upload(xt._ctx, C.c_void_p(xdev + i * xfb), fr.ctypes.data, xfb))
                xt.set_reference(xst.frame(0), xst.ref_xy)
                xptrs = (C.c_void_p * xf)(*[xdev + ((1 + j) % xring) * xfb for j in range(xf)])
                xout = (N.Record * xf)()

                def xstep():
                    xt.track_device_ptrs(xptrs, 1, xout)

                xt.set_kernel_timing(True)
                xms, _, xs = timed(xstep, 3, 2, xt)
                xmacs = xs.ncc_algorithmic_macs / max(xs.ncc_kernel_launches, 1)
                xk = xs.ncc_kernel_ms / max(xs.ncc_kernel_launches, 1)
                extra["config_" + name] = {
                    "config": config_dict(name, xcfg), "value": xf * 3 / (xms * 1e-3), "unit": UNIT, "ms_per_step": xms / 3,
                    "frames_per_step": xf, "lanes": 1, "accepted_frames_last_step": sum(1 for r in xout if r.status == 0),
                    "match_top_s": (2.0 * xmacs / (xk * 1e-3) / 1e12) if xk > 0 else None,
                    "match_frac_of_int8_tensor_peak": (2.0 * xmacs / (xk * 1e-3) / 1e12 / imma_peak) if (xk > 0 and imma_peak) else None,
                    "kernel_ms_per_step": {"lst_preprocess": xs.preprocess_ms / 3, "match": xs.ncc_kernel_ms / 3, "lst_epilogue": xs.epilogue_ms / 3},
                    "preprocess_tflops": (38.0 * xs.preprocess_px / (xs.preprocess_ms * 1e-3) / 1e12) if xs.preprocess_ms > 0 else None,
                    "preprocess_hbm_gbs": (xs.preprocess_px * ((xbits // 8) + 1) / (xs.preprocess_ms * 1e-3) / 1e9) if xs.preprocess_ms > 0 else None,
                }
                # several lanes where the step is long enough to give each of them full batches (config A): the host-side chain
                # walk of one lane runs under the kernels of the others, as in the main leg
                xlanes = max(1, min(LANES_DEFAULT, len(my_cores)))
                if xcfg.s_area > 0 and xf // xmb >= 4 and xlanes > 1:
                    xm = LaserSpotTrack4Multi(xcfg.width, xcfg.height, xbits, [local_rank] * xlanes, templSize=xcfg.templ_size,
                                              sArea=xcfg.s_area, max_batch=xmb)
                    xm.set_reference(xst.frame(0), xst.ref_xy)

                    def xstep_multi():
                        xm.track_device_ptrs(xptrs, 1, xout)

                    xms2, _, _ = timed(xstep_multi, 3, 2, xm)
                    e = extra["config_" + name]
                    e["value_one_lane"], e["ms_per_step_one_lane"] = e["value"], e["ms_per_step"]
                    e["value"], e["ms_per_step"], e["lanes"] = xf * 3 / (xms2 * 1e-3), xms2 / 3, xlanes
                    e["accepted_frames_last_step"] = sum(1 for r in xout if r.status == 0)
                    xm.close()
                xt.device_free(xdev)
                xt.close()
            except Exception as e:      # the extras never fail the run
                extra["config_" + name] = {"error": str(e)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample or sample_d
        rate, n, ok = cpu_reference_rate(args.config, ring, sample, 1)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": f"first {n} frames of the same ring, oracle loop (numpy ImageJ blur + "
                                  f"cv2.matchTemplate TM_CCOEFF_NORMED, IPP off), 1 thread"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": config_dict(args.config, cfg),
            "details": {"frames_per_step_per_gpu": F, "ring_frames": ring, "ring_bytes": ring * fb,
                        "l2": "inputs larger than L2 (ring of distinct slices cycled)" if ring * fb > 126e6 else
                              "ring smaller than L2; windows are re-read from HBM/L2 each pass",
                        "max_batch": args.max_batch or mb_d,
                        "parallelism": f"frames sharded over {world} GPU(s), no collective" + (f"; {lanes} lanes per GPU" if lanes > 1 else ""),
                        "lanes_per_gpu": lanes, "host_cores_of_this_rank": len(my_cores),
                        "accepted_frames_last_step": ok_frames},
            "clocks": clocks, "e2e": e2e, "e2e_files": e2e_files, "e2e_jpeg": e2e_jpeg, "e2e_pageable": e2e_pageable, "gpu_launches": int(s_dev.kernel_launches),
            "roofline": roofline, "roofline_idp4a_engine": roofline_alt, "cpu_baseline": cpu_baseline, "extra": extra,
            "resolver": {"passes": int(s_dev.passes), "tasks": int(s_dev.tasks), "respeculated": int(s_dev.respeculated),
                         "chunks_retracked": retr_all[0], "chunks_retracked_late": getattr(sh_dev, "retracked_late", 0),
                         "lane_chunks_retracked": retr_all[1], "continuous_stack": bool(args.continuous), "lead_in": LEAD,
                         "boundary_exchange": ("pipelined: overlaps the next step" if pipelined else "synchronous") +
                                              (", " + os.environ.get("LST_BENCH_EXCHANGE", "nccl") if world > 1 else ""), "host_ms_per_step": s_dev.host_ms / K,
                         "device_wait_ms_per_step": s_dev.device_wait_ms / K},
            "wall_ms_per_step": wall_dev / K,
        }
        print(json.dumps(line), flush=True)
    if runner is not trk:
        runner.close()
    trk.device_free(dev_ring)
    lib.lst_free_pinned(host_ring)
    trk.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries the JSON line and nothing else: whatever libraries write to file descriptor 1 meanwhile (NCCL's version
    # banner, for one) goes to stderr
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, "w", buffering=1)
    main()
