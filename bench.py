#!/usr/bin/env python
"""Benchmark of the video-to-voxel hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE config 2 — the train_v2v_e2vid_10k data
path: WebVid-shaped synthetic clips uint8 [121,480,640], num_bins=5,
frames_per_bin=1, thresholds / noise parameters sampled per clip by the
reference's law (data/v2v_datasets.py:368-386) with the shipped ranges
(config/train_v2v_e2vid_10k.yaml:72-75), noise generated in-kernel (Philox).
A step = one pass of the fused kernel over one batch of clips per GPU.

* value      : whole-job Mpix-frames/s (pixel-intervals/s / 1e6), inputs resident in HBM, CUDA-event timed.
* e2e        : same metric through the host-buffer API (pinned uint8 frames in, float32 voxels back in
               pinned host memory; H2D + kernel + D2H inside the timed region).
* roofline   : algorithmic bytes (1 B in + 4 B out per pixel-interval, +1 B/pixel for frame 0) / launch time
               against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* cpu_baseline: the NumPy oracle port of the reference's ESIM core timed on this host's cores.

Under torchrun (N>1) every rank simulates its own clips (clip-sharded, weak scaling, no data-path
collective); one NCCL all-reduce of the event-count statistics closes the timed region.
"""
from __future__ import annotations

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

N_FRAMES, H, W, BINS, FPB = 121, 480, 640, 5, 1
TRAIN_CFG = dict(num_bins=BINS, frames_per_bin=FPB, threshold_range=[0.05, 2], max_thres_pos_neg_gap=1.5,
                 base_noise_std_range=[0, 0.1], hot_pixel_fraction_range=[0, 0.001], hot_pixel_std_range=[0, 10])
PIX_INTERVALS_PER_CLIP = (N_FRAMES - 1) * H * W
ALGO_BYTES_PER_CLIP = H * W * (N_FRAMES * 1 + (N_FRAMES - 1) // FPB * 4)       # 184.6 MB


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/), else None."""
    p = os.path.join(ROOT, "profiles", "esim_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# CPU baseline: NumPy port of the reference's ESIM core (oracle/), one clip per worker process
# ----------------------------------------------------------------------------------------------

def find_reference():
    """The reference checkout, if this machine has one (the build container does, the GPU box does not): $V2V_REFERENCE,
    baseline/_ref, /root/reference.  Returns its path or None."""
    for p in (os.environ.get("V2V_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if p and os.path.isfile(os.path.join(p, "data", "v2v_core_esim.py")):
            return p
    return None


def _reference_emulator(ref):
    """data/v2v_core_esim.py imports only numpy: load the reference's own EventEmulator from its file."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_v2v_core_esim", os.path.join(ref, "data", "v2v_core_esim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.EventEmulator


def _cpu_clip_job(args):
    seed, n_frames, h, w = args[:4]
    ref = args[4] if len(args) > 4 else None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import v2v_oracle as orc
    g = np.random.Generator(np.random.PCG64(seed))
    base = g.integers(0, 256, size=(h, w)).astype(np.int16)
    steps = g.integers(-6, 7, size=(n_frames, h, w), dtype=np.int16)
    steps[0] = 0
    vid = np.clip(base[None] + np.cumsum(steps, axis=0, dtype=np.int16), 0, 255).astype(np.uint8)
    rs = np.random.RandomState(seed)
    p = orc.sample_esim_params(rs, TRAIN_CFG["threshold_range"], TRAIN_CFG["max_thres_pos_neg_gap"],
                               TRAIN_CFG["base_noise_std_range"], TRAIN_CFG["hot_pixel_fraction_range"],
                               TRAIN_CFG["hot_pixel_std_range"])
    t0 = time.perf_counter()
    # the timed part is what the reference's imgs_to_voxels does: draws + simulation + binning + float32 cast
    if ref is not None:           # the reference's own code (data/v2v_core_esim.py:26-69, data/v2v_datasets.py:388-400)
        np.random.seed(seed)
        iv = _reference_emulator(ref)(pos_thres=p["pos_thres"], neg_thres=p["neg_thres"], base_noise_std=p["base_noise_std"],
                                      hot_pixel_fraction=p["hot_pixel_fraction"], hot_pixel_std=p["hot_pixel_std"],
                                      put_noise_external=False, seed=None).video_to_voxel(vid)
        vox = iv.reshape(((n_frames - 1) // (BINS * FPB), BINS, FPB, h, w)).sum(axis=2).astype(np.float32)
    else:
        u0, hot, gs = orc.esim_draw_randomness(n_frames, h, w, p["hot_pixel_fraction"], p["hot_pixel_std"], rs)
        iv = orc.esim_video_to_voxel(vid, p["pos_thres"], p["neg_thres"], p["base_noise_std"], u0, hot, gs, False)
        vox = orc.bin_accumulate(iv, BINS, FPB).astype(np.float32)
    dt = time.perf_counter() - t0
    return dt, float(np.abs(vox).sum())


def cpu_reference_run(n_frames, clips, procs, ref=None):
    """Simulate `clips` clips of [n_frames,H,W] on `procs` processes; returns wall seconds."""
    import multiprocessing as mp
    jobs = [(1000 + i, n_frames, H, W, ref) for i in range(clips)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_clip_job(j) for j in jobs]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            pool.map(_cpu_clip_job, [(0, 6, 8, 8)] * procs)        # warm the workers (imports) outside the timing
            t0 = time.perf_counter()
            res = pool.map(_cpu_clip_job, jobs)
    wall = time.perf_counter() - t0
    return wall, res


CPU_IMPL = {None: "NumPy oracle port of data/v2v_core_esim.py (oracle/v2v_oracle.py; bit-identical to the reference by "
                  "tests/golden/make_golden.py) incl. MT19937 noise draws, binning, float32 cast"}


def cpu_impl_name(ref):
    return CPU_IMPL[None] if ref is None else (f"the reference's own EventEmulator.video_to_voxel ({ref}/data/v2v_core_esim.py) "
                                               "+ the bin sum and float32 cast of data/v2v_datasets.py:399-400,340-349")


def cpu_baseline(cores):
    ref = find_reference()
    n_frames = N_FRAMES                                             # full config-2 clips, one per host core
    wall, _ = cpu_reference_run(n_frames, cores, cores, ref)
    pix = cores * (n_frames - 1) * H * W
    return {"value": pix / wall / 1e6, "unit": "Mpix-frames/s", "cores": cores, "kind": "reference" if ref else "port",
            "sample": f"{cores} clips uint8 [{n_frames},{H},{W}] (full config-2 clips), one per process: " + cpu_impl_name(ref),
            "clips_per_s_equiv": pix / wall / PIX_INTERVALS_PER_CLIP, "wall_s": wall}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the same workload on all host cores (rank 0 only).
    A step is a bounded sample of the workload — full [121,480,640] clips, one per host core — so that K+W steps end
    within a few minutes; throughput is normalised per pixel-interval, so the sample size does not bias it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ref = find_reference()
    import multiprocessing as mp
    jobs_per_step = cores
    n_frames = N_FRAMES
    times = []
    budget_s = float(os.environ.get("V2V_REF_BUDGET_S", "240"))
    with mp.get_context("spawn").Pool(cores) as pool:
        pool.map(_cpu_clip_job, [(0, 6, 8, 8, ref)] * cores)
        # calibrate on one full-length step (it is also the first warm-up step): if K+W such steps would not fit the
        # budget, keep the clips per step (= cores, the reference's DataLoader-worker parallelism) and shorten the clips
        t0 = time.perf_counter()
        pool.map(_cpu_clip_job, [(4000 + i, n_frames, H, W, ref) for i in range(jobs_per_step)])
        t_full = time.perf_counter() - t0
        total_steps = max(1, args.warmup - 1) + args.steps
        if t_full * total_steps > budget_s:
            n_frames = int(max(BINS * FPB * 2 + 1, ((N_FRAMES - 1) * budget_s / (t_full * total_steps)) // (BINS * FPB) * (BINS * FPB) + 1))
        for s in range(max(0, args.warmup - 1) + args.steps):
            jobs = [(5000 + s * jobs_per_step + i, n_frames, H, W, ref) for i in range(jobs_per_step)]
            t0 = time.perf_counter()
            pool.map(_cpu_clip_job, jobs)
            if s >= max(0, args.warmup - 1):
                times.append(time.perf_counter() - t0)
    total = sum(times)
    pix = args.steps * jobs_per_step * (n_frames - 1) * H * W
    val = pix / total / 1e6
    sample = (f"{args.steps} steps x {jobs_per_step} clips uint8 [{n_frames},{H},{W}]"
              + ("" if n_frames == N_FRAMES else f" (clips shortened from {N_FRAMES} frames to fit {budget_s:.0f} s)")
              + ", one clip per host process: " + cpu_impl_name(ref))
    line = {
        "impl": "reference", "metric": "video_to_voxel_throughput", "value": val, "unit": "Mpix-frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "clips_per_s": pix / total / PIX_INTERVALS_PER_CLIP,
        "config": workload_config(args.clips),
        "cpu_baseline": {"value": val, "unit": "Mpix-frames/s", "cores": cores, "kind": "reference" if ref else "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(clips_per_step, note=""):
    return {"workload": "BASELINE config 2: train_v2v_e2vid_10k data path, synthetic WebVid-shaped clips uint8 "
                        f"[{N_FRAMES},{H},{W}], num_bins={BINS}, frames_per_bin={FPB}, per-clip thresholds U[0.05,2]x gap U[1,1.5], "
                        "base_noise_std U[0,0.1], hot_pixel_fraction U[0,0.001], hot_pixel_std U[0,10], in-kernel Philox noise",
            "clips_per_step_per_gpu": clips_per_step, "frames": N_FRAMES, "height": H, "width": W, "num_bins": BINS,
            "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)", "note": note,
            "launch": "timed steps replayed from one CUDA graph (each step = one C-ABI kernel launch + stats reduction)"}


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock / throttle reasons during the timed region, sampled in-process through NVML every `period` s
    (a polling nvidia-smi subprocess takes driver locks that stall kernel launches; it is only the fallback)."""

    def __init__(self, index, period=0.1):
        self.period = period
        self.index, self.rows, self.stop_flag, self.thread, self.h = index, [], False, None, None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self, t0, t1):
        if self.nv is None:
            return self._smi_once()
        time.sleep(0.05)
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        inside = [(sm, rs) for t, sm, rs in self.rows if t0 <= t <= t1] or [(sm, rs) for _, sm, rs in self.rows[-3:]]
        reasons = sorted({n for _, rs in inside for n, bit in names.items() if rs & bit})
        sms = [sm for sm, _ in inside]
        return {"sm_mhz": float(np.median(sms)) if sms else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sms), "source": "nvml"}

    def _smi_once(self):
        try:
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            f = [x.strip() for x in out]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6])
                       if v.lower().startswith("active")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------

def make_clips(torch, dev, clips, seed):
    """Temporally correlated random-walk clips generated on the device (SURVEY §8(d) synthetic input)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((clips, N_FRAMES, H, W), dtype=torch.uint8, device=dev)
    for b in range(clips):
        base = torch.randint(0, 256, (H, W), generator=g, device=dev, dtype=torch.int16)
        steps = torch.randint(-6, 7, (N_FRAMES, H, W), generator=g, device=dev, dtype=torch.int16)
        steps[0] = 0
        out[b] = (base[None] + torch.cumsum(steps, dim=0)).clamp_(0, 255).to(torch.uint8)
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cores)                                  # before CUDA is initialised in this process

    import torch
    import v2v_b200 as v2v
    from v2v_b200 import dist as vdist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = "unchanged"
    try:        # run on the CPUs next to this GPU so that pinned staging buffers land on its NUMA node (PCIe locality)
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = f"pinned to {len(os.sched_getaffinity(0))} GPU-local cpus"
    except Exception as e:      # not fatal: only affects host<->device copy locality
        numa = f"unchanged ({type(e).__name__})"
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set on the host: keep stdout to the one JSON line
        # by pointing fd 1 at stderr while the communicator is created (first collective)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = args.clips
    vz = v2v.V2VVoxelizer(TRAIN_CFG, device=dev)
    # weak scaling: every rank gets the same synthetic clips and sampled parameters (identical work per GPU, so the
    # per-N numbers compare hardware, not the luck of the draw: the cost of a clip depends on its thresholds — a rank
    # with different random parameters ran 8 % slower); the noise streams differ through the global clip index
    rs = np.random.RandomState(1234)
    params = vz.sample_batch_params(B, rs=rs)
    frames = make_clips(torch, dev, B, 100)
    T = (N_FRAMES - 1) // (BINS * FPB)
    out = torch.empty((B, T, BINS, H, W), dtype=torch.float32, device=dev)
    col = lambda k: torch.tensor([p[k] for p in params], dtype=torch.float64, device=dev)
    pos, neg, std, frac, hstd = (col(k) for k in ("pos_thres", "neg_thres", "base_noise_std", "hot_pixel_fraction", "hot_pixel_std"))
    stats_total = torch.zeros(2, dtype=torch.int64, device=dev)

    def step(i):
        o = v2v.frames_to_voxel(frames, pos, neg, num_bins=BINS, frames_per_bin=FPB, noise="philox", base_noise_std=std,
                                hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=args.seed, clip_index_base=(i * world + rank) * B,
                                with_stats=True, out=out)
        return o

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i)
    barrier()
    # The K timed steps are captured once in CUDA graphs of `gsteps` steps and replayed, so the GPU runs them back
    # to back whatever the host is doing (NVML sampling, other ranks' processes).  Every captured step is the full
    # public call: kernel launch through the C ABI + per-step stats reduction.
    gsteps = max(1, min(args.graph_steps, args.steps))
    while args.steps % gsteps:
        gsteps -= 1
    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(2):
                stats_total += step(args.warmup + i).stats.sum(dim=0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        stats_total.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(gsteps):
                o = step(args.warmup + i)
                stats_total += o.stats.sum(dim=0)
        graph.replay()                                             # one untimed replay
        torch.cuda.synchronize(dev)
        stats_total.zero_()
    sampler = ClockSampler(local, args.clock_period)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    launches0 = v2v.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    torch.cuda.nvtx.range_push("timed")
    ev0.record()
    if graph is not None:
        for _ in range(args.steps // gsteps):
            graph.replay()
    else:
        for i in range(args.steps):
            o = step(args.warmup + i)
            stats_total += o.stats.sum(dim=0)
    ev_mid = torch.cuda.Event(enable_timing=True)
    ev_mid.record()                                                # this rank's own steps end here
    job = vdist.pack_stats(stats_total.view(1, 2), args.steps * B * PIX_INTERVALS_PER_CLIP, args.steps * B, device=dev)
    vdist.allreduce_stats(job)                                     # the only collective: event-count statistics (NCCL)
    ev1.record()
    barrier()
    torch.cuda.nvtx.range_pop()
    t_wall1 = time.perf_counter()
    launches = (v2v.launch_count() - launches0) if graph is None else args.steps      # one kernel of ours per step (graph replays are not re-counted by the library)
    ms = ev0.elapsed_time(ev1)
    own_ms = ev0.elapsed_time(ev_mid)
    per_rank = [own_ms / args.steps]
    if dist is not None:
        t = torch.tensor([own_ms], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [float(x.item()) / args.steps for x in allt]   # each rank's own steps, before the stats all-reduce
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = (sampler.stop(t_wall0, t_wall1) if not args.no_clocks else sampler._smi_once()) if rank == 0 else None

    # kernel-only duration for the roofline: per-launch CUDA events on the launching stream
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i, (a, b) in enumerate(evs):
        a.record()
        v2v.frames_to_voxel(frames, pos, neg, num_bins=BINS, frames_per_bin=FPB, noise="philox", base_noise_std=std,
                            hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=args.seed, clip_index_base=i * B, out=out,
                            with_stats=True)            # the very kernel variant the timed loop launches
        b.record()
    torch.cuda.synchronize(dev)
    launch_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # noise-free variant (explains how much of the time is the in-kernel generator)
    evs2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(3, min(args.steps // 2, 20)))]
    v2v.frames_to_voxel(frames, pos, neg, num_bins=BINS, frames_per_bin=FPB, noise="none", out=out)   # first launch of this variant: untimed
    torch.cuda.synchronize(dev)
    for a, b in evs2:
        a.record()
        v2v.frames_to_voxel(frames, pos, neg, num_bins=BINS, frames_per_bin=FPB, noise="none", out=out)
        b.record()
    torch.cuda.synchronize(dev)
    clean_ms = float(np.mean([a.elapsed_time(b) for a, b in evs2]))

    # ---- e2e: pinned host frames -> H2D -> kernel -> D2H float32 voxels in pinned host memory ----
    e2e = None
    if not args.no_e2e:
        Be = min(B, args.e2e_clips)
        host_in = torch.empty((Be, N_FRAMES, H, W), dtype=torch.uint8).pin_memory()
        host_in.copy_(frames[:Be].cpu())
        host_out = torch.empty((Be, T, BINS, H, W), dtype=torch.float32).pin_memory()
        pipe = v2v.HostPipeline(vz, dev, clips_per_chunk=args.e2e_chunk, seed=args.seed)
        ksteps = max(2, min(args.steps, 5))
        pipe.run(host_in, params[:Be], host_out)                   # warm-up (allocates staging buffers)
        barrier()
        t0 = time.perf_counter()
        for i in range(ksteps):
            pipe.run(host_in, params[:Be], host_out)
        torch.cuda.synchronize(dev)
        t_e2e = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        e2e = {"value": world * ksteps * Be * PIX_INTERVALS_PER_CLIP / t_e2e / 1e6, "unit": "Mpix-frames/s",
               "h2d_bytes_per_step": int(host_in.numel()), "d2h_bytes_per_step": int(host_out.numel() * 4),
               "clips_per_s": world * ksteps * Be / t_e2e, "steps": ksteps, "clips_per_step_per_gpu": Be,
               "api": "v2v_b200.HostPipeline.run(pinned uint8 frames, params, pinned float32 out): chunked H2D / kernel / D2H on 3 streams",
               "cpu_affinity": numa}
        # the training path keeps the voxels on the GPU for the model (train.py:79-83): host frames in, stats out
        pipe.run(host_in, params[:Be], None)
        barrier()
        t0 = time.perf_counter()
        for i in range(ksteps):
            st = pipe.run(host_in, params[:Be], None)
            st_host = st.sum(dim=0).cpu()                          # D2H read of the step's statistics (16 bytes)
        torch.cuda.synchronize(dev)
        t_tr = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([t_tr], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_tr = float(t.item())
        e2e["device_resident_output"] = {"value": world * ksteps * Be * PIX_INTERVALS_PER_CLIP / t_tr / 1e6, "unit": "Mpix-frames/s",
                                         "clips_per_s": world * ksteps * Be / t_tr, "h2d_bytes_per_step": int(host_in.numel()),
                                         "d2h_bytes_per_step": 16,
                                         "note": "same pipeline without the voxel D2H: what the train loop sees (voxels consumed on the GPU)"}

    if e2e is not None:
        # what the host link allows: pinned H2D and D2H copies of 1 GiB, best of 3 (PCIe Gen5 x16 on the HGX boards)
        hb = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        db = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        link = {}
        for name, (dst, src) in (("h2d", (db, hb)), ("d2h", (hb, db))):
            best = 0.0
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                a.record()
                dst.copy_(src, non_blocking=True)
                b.record()
                torch.cuda.synchronize(dev)
                best = max(best, (1 << 30) / (a.elapsed_time(b) * 1e-3) / 1e9)
            link[name + "_gbs"] = best
        # ... and the D2H rate while H2D traffic in the pipeline's own proportion (1 byte in per 4 bytes out) runs beside it
        hb2 = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
        db2 = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
        s_h2d = torch.cuda.Stream(dev)
        best = 0.0
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s_h2d.wait_stream(torch.cuda.current_stream(dev))
            a.record()
            with torch.cuda.stream(s_h2d):
                db2.copy_(hb2, non_blocking=True)
            hb.copy_(db, non_blocking=True)
            b.record()
            torch.cuda.synchronize(dev)
            best = max(best, (1 << 30) / (a.elapsed_time(b) * 1e-3) / 1e9)
        link["d2h_gbs_with_h2d_beside"] = best
        del hb, db, hb2, db2
        per_clip_s = max(N_FRAMES * H * W / (link["h2d_gbs"] * 1e9), T * BINS * H * W * 4 / (link["d2h_gbs"] * 1e9))
        link["ceiling_duplex_clips_per_s_per_gpu"] = link["d2h_gbs_with_h2d_beside"] * 1e9 / (T * BINS * H * W * 4)
        link["ceiling_clips_per_s_per_gpu"] = 1.0 / per_clip_s
        link["note"] = ("measured on this rank while all ranks copy at the same time; ceiling = 1 / max(frame bytes / h2d, voxel bytes / d2h) "
                        "per clip from the one-directional rates (the float32 voxels are 80 % of the bytes); ceiling_duplex = the same "
                        "with the D2H rate measured while H2D traffic in the pipeline's proportion runs beside it: what the link gives "
                        "this byte mix")
        e2e["host_link"] = link
        e2e["frac_of_host_link_ceiling"] = e2e["clips_per_s"] / world / link["ceiling_clips_per_s_per_gpu"]
        e2e["frac_of_duplex_ceiling"] = e2e["clips_per_s"] / world / link["ceiling_duplex_clips_per_s_per_gpu"]

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        del out, frames
        torch.cuda.empty_cache()
        from tools.bench_configs import secondary_configs
        secondary = secondary_configs(dev, cpu_arm=not args.no_cpu_baseline)

    if rank == 0:
        peak, peak_src = measured_peak()
        total_pix = world * args.steps * B * PIX_INTERVALS_PER_CLIP
        achieved = B * ALGO_BYTES_PER_CLIP / (launch_ms * 1e-3) / 1e9
        tr = ncu_traffic()
        line = {
            "metric": "video_to_voxel_throughput", "value": total_pix / (ms * 1e-3) / 1e6, "unit": "Mpix-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(B),
            "clips_per_s": world * args.steps * B / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (tr or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "kernel": "esim_fast_kernel<PHILOX, stats> (v2v_b200/csrc/esim_fast.cu)", "algorithmic_bytes_per_launch": B * ALGO_BYTES_PER_CLIP,
                         "launch_ms": launch_ms, "frac_of_8TBs_nominal": achieved / 8000.0,
                         "noise_free_launch_ms": clean_ms,
                         "noise_free_frac": B * ALGO_BYTES_PER_CLIP / (clean_ms * 1e-3) / 1e9 / peak},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "ms_per_step_per_rank": per_rank,
            "event_stats": vdist.stats_dict(job),
            "secondary": secondary,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_config5(args):
    """BASELINE config 5: clip-sharded 1080p 5-bin voxel generation feeding an E2VID forward, 10 k clips on the GPUs of one box.
    Clip i of the job goes to rank i % world (v2v_b200.dist.shard_indices, the DistributedSampler rule); a step is one
    batch of clips per GPU: frames [b,26,1080,1920] (a resident pool of synthetic clips, re-used round robin) -> fused
    kernel writing voxels straight into the /16-padded consumer layout [b,5,5,1088,1920] + ground-truth frames + stats ->
    E2VID-shaped recurrent U-Net forward over the 5 voxels on the same stream (tools/e2vid_consumer.py, cuDNN).  `--steps`
    steps are timed per rank (the whole 10 k-clip job is `--c5-clips / (world * --c5-batch)` such steps); one NCCL
    all-reduce of the statistics vector closes the timed region."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import v2v_b200 as v2v
    from v2v_b200 import dist as vdist
    from tools.e2vid_consumer import E2VIDShaped
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n5, h5, w5, b5 = 26, 1080, 1920, args.c5_batch
    mine = vdist.shard_indices(args.c5_clips, rank, world)
    vz = v2v.V2VVoxelizer(TRAIN_CFG, device=dev)
    params = vz.sample_batch_params(b5, rs=np.random.RandomState(99))
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    pool = torch.empty((2 * b5, n5, h5, w5), dtype=torch.uint8, device=dev)          # resident pool of clips
    for b in range(2 * b5):
        base = torch.randint(0, 256, (h5, w5), generator=g, device=dev, dtype=torch.int16)
        st = torch.randint(-6, 7, (n5, h5, w5), generator=g, device=dev, dtype=torch.int16)
        st[0] = 0
        pool[b] = (base[None] + torch.cumsum(st, 0)).clamp_(0, 255).to(torch.uint8)
    store = torch.zeros((b5, 5, 5, 1088, 1920), dtype=torch.float32, device=dev)     # pads written once, here
    net = E2VIDShaped().to(dev).eval()
    stats_total = torch.zeros(2, dtype=torch.int64, device=dev)

    def sim(i):
        fr = pool[(i % 2) * b5:(i % 2 + 1) * b5]
        return vz.batch_to_tensors(fr, params, seed=args.seed, clip_index_base=mine[(i * b5) % max(len(mine), 1)] if mine else 0,
                                   pad_multiple=16, with_stats=True, out=store)

    def step(i, with_model=True):
        o = sim(i)
        stats_total.add_(o["stats"].sum(dim=0))
        if with_model:
            return net.forward_sequence(o["events_padded"])
        return None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(k, with_model):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(k):
            step(i, with_model)
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(3, args.warmup)):
        step(i)
    sampler = ClockSampler(local, args.clock_period)
    if rank == 0:
        sampler.start()
    stats_total.zero_()
    t0 = time.perf_counter()
    ms = timed(args.steps, True)
    job = vdist.pack_stats(stats_total.view(1, 2), args.steps * b5 * 25 * h5 * w5, args.steps * b5, device=dev)
    vdist.allreduce_stats(job)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_sim = timed(args.steps, False)                      # the simulator alone (same launches, no consumer)
    # e2e: every step's clips come from pinned host memory (H2D inside the timed region) and a result of the step goes back
    # (the mean of the last reconstructed frame of every clip, read on the host)
    host_pool = pool.cpu().pin_memory()
    dev_in = torch.empty((b5, n5, h5, w5), dtype=torch.uint8, device=dev)
    host_res = torch.empty((b5,), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        dev_in.copy_(host_pool[(i % 2) * b5:(i % 2 + 1) * b5], non_blocking=True)
        o = vz.batch_to_tensors(dev_in, params, seed=args.seed, clip_index_base=mine[(i * b5) % max(len(mine), 1)] if mine else 0,
                                pad_multiple=16, with_stats=True, out=store)
        rec = net.forward_sequence(o["events_padded"])                 # [B,T,1,Hp,Wp]
        host_res.copy_(rec[:, -1].reshape(b5, -1).mean(dim=1), non_blocking=True)

    e2e_step(0)
    barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for i in range(args.steps):
        e2e_step(i)
    eb.record()
    barrier()
    ms_e2e = ea.elapsed_time(eb)
    if dist is not None:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i, (a, b) in enumerate(evs):
        a.record()
        sim(i)
        b.record()
    torch.cuda.synchronize(dev)
    launch_ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    if rank == 0:
        peak, peak_src = measured_peak()
        by = b5 * h5 * w5 * (n5 + 25 * 4 + 5 * 4)          # frames in, voxels out, ground-truth frames out
        clips_s = world * args.steps * b5 / (ms * 1e-3)
        line = {
            "metric": "video_to_voxel_throughput", "value": world * args.steps * b5 * 25 * h5 * w5 / (ms * 1e-3) / 1e6, "unit": "Mpix-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (simulator) / f32 (consumer)", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: clip-sharded 1080p 5-bin voxel generation feeding an E2VID-shaped forward "
                                   f"({args.c5_clips} clips sharded rank::world; clips uint8 [26,1080,1920] -> 5 voxels of 5 bins in the /16-padded "
                                   "layout [5,5,1088,1920] + 5 ground-truth frames + statistics, in-kernel Philox noise, then the recurrent U-Net "
                                   "forward over the 5 voxels, cuDNN)",
                       "clips_per_step_per_gpu": b5, "frames": n5, "height": h5, "width": w5, "num_bins": 5, "job_clips": args.c5_clips,
                       "job_steps_per_gpu": -(-len(mine) // b5), "l2": "each step moves 1.3 GB of frames and voxels (> 126 MB L2)"},
            "clips_per_s": clips_s, "job_seconds_estimate": args.c5_clips / clips_s,
            "simulator_only": {"ms_per_step": ms_sim / args.steps, "clips_per_s": world * args.steps * b5 / (ms_sim * 1e-3),
                               "share_of_step": ms_sim / ms},
            "roofline": {"bound": "hbm", "achieved": by / (launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": by / (launch_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "esim_fast_kernel<PHILOX, frames, stats> (v2v_b200/csrc/esim_fast.cu)", "launch_ms": launch_ms,
                         "algorithmic_bytes_per_launch": by},
            "e2e": {"value": world * args.steps * b5 * 25 * h5 * w5 / (ms_e2e * 1e-3) / 1e6, "unit": "Mpix-frames/s",
                    "clips_per_s": world * args.steps * b5 / (ms_e2e * 1e-3), "h2d_bytes_per_step": b5 * n5 * h5 * w5,
                    "d2h_bytes_per_step": b5 * 4,
                    "api": "pinned uint8 clips -> H2D -> V2VVoxelizer.batch_to_tensors (padded layout) -> E2VID-shaped forward -> "
                           "per-clip mean of the last reconstruction read on the host"},
            "cpu_baseline": None,
            "gpu_launches": args.steps, "clocks": clocks, "event_stats": vdist.stats_dict(job),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 100; 20 for --impl reference, so its clips keep their full length)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=32, help="clips per step per GPU")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--e2e-clips", type=int, default=16)
    ap.add_argument("--e2e-chunk", type=int, default=1, help="clips per H2D / kernel / D2H chunk of the host pipeline (same-box sweep: 1: 368, 2: 363, 4: 353 clips/s)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary block (BASELINE configs 1, 3, 4, 5-shape, train batch)")
    ap.add_argument("--config", default="c2", choices=["c2", "c5"],
                    help="c2: BASELINE config 2 (the headline); c5: config 5, clip-sharded 1080p voxel generation feeding an E2VID-shaped forward")
    ap.add_argument("--c5-clips", type=int, default=10000, help="config 5: clips of the whole job (sharded rank::world)")
    ap.add_argument("--c5-batch", type=int, default=4, help="config 5: clips per step per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch every timed step from Python instead of replaying CUDA graphs")
    ap.add_argument("--graph-steps", type=int, default=100)
    ap.add_argument("--no-clocks", action="store_true", help="(experiments) do not sample clocks during the timed region")
    ap.add_argument("--clock-period", type=float, default=0.1)
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else (10 if args.config == "c5" else 100)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config == "c5":
        run_config5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
