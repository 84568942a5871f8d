#!/usr/bin/env python
"""Secondary BASELINE configs (1, 3, 4, 5-shape) timed on one GPU: device-resident inputs, CUDA events.

    python tools/bench_configs.py [--out profiles/configs_rNN.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import v2v_b200 as v2v  # noqa: E402
from v2v_b200.v2e import frames_to_voxel_v2e  # noqa: E402

PEAK = 6549.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms)), float(np.min(ms))


def timeit_graph(fn, n=20, reps=5):
    """The same call captured n times in one CUDA graph and replayed: what a training loop that captures its data path
    sees (no per-call Python / ctypes / allocator time between the launches).  Returns ms per call (min over replays)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b) / n)
    return float(np.min(ms))


def walk(B, N, H, W, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty((B, N, H, W), dtype=torch.uint8, device=dev)
    for b in range(B):
        base = torch.randint(0, 256, (H, W), generator=g, device=dev, dtype=torch.int16)
        st = torch.randint(-6, 7, (N, H, W), generator=g, device=dev, dtype=torch.int16)
        st[0] = 0
        out[b] = (base[None] + torch.cumsum(st, 0)).clamp_(0, 255).to(torch.uint8)
    return out


def _cpu_arm(fn, what):
    """One bounded CPU run of the oracle port of the same path (bench.py's secondary block; ~1-3 s each)."""
    t0 = time.perf_counter()
    units = fn()
    dt = time.perf_counter() - t0
    return {"seconds": dt, "units": units, "sample": what, "kind": "port", "cores": 1}


def secondary_configs(dev, cpu_arm=True):
    """BASELINE configs 1, 3, 4, the config-5 shape and the shipped training batch on one GPU (device-resident inputs,
    CUDA events, median of the iterations), each with its algorithmic bytes, fraction of the measured HBM peak and — when
    `cpu_arm` — the NumPy / C oracle port of the same path timed on one host core on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import v2v_oracle as orc
    res = {"peak_gbs": PEAK, "gpu": torch.cuda.get_device_name(dev)}

    # config 1: one 40x256x256 clip, raw core (num_bins=1), fixed thresholds, no noise: latency bound
    fr = walk(1, 40, 256, 256, 1, dev)
    u0 = torch.rand((1, 256, 256), dtype=torch.float64, device=dev)
    out1 = torch.empty((1, 39, 1, 256, 256), dtype=torch.float32, device=dev)
    pos = torch.full((1,), 0.2, dtype=torch.float64, device=dev)
    f1 = lambda: v2v.frames_to_voxel(fr, pos, pos, num_bins=1, u0=u0, out=out1)
    med, mn = timeit(f1, 20)
    by = 40 * 65536 + 39 * 65536 * 4
    res["config1_one_clip_40x256x256"] = {"ms": med, "ms_min": mn, "ms_graph_replay": timeit_graph(f1), "clips_per_s": 1e3 / med, "Mpix_frames_per_s": 39 * 65536 / med / 1e3,
                                          "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK, "note": "single small clip: launch/latency bound (SURVEY §7)"}
    # ... and as the drop-in call a user of the reference makes: EventEmulator(...).video_to_voxel, NumPy in -> NumPy out
    vid1 = fr[0].cpu().numpy()
    em = v2v.EventEmulator(pos_thres=0.2, neg_thres=0.2, base_noise_std=0.0, hot_pixel_fraction=0.0, hot_pixel_std=0.0, rng="philox",
                           seed=1, device=dev)
    em.video_to_voxel(vid1)
    t0 = time.perf_counter()
    for _ in range(10):
        em.video_to_voxel(vid1)
    res["config1_one_clip_40x256x256"]["dropin_numpy_in_out_ms"] = (time.perf_counter() - t0) * 100
    if cpu_arm:
        c = _cpu_arm(lambda: (orc.esim_video_to_voxel(vid1, 0.2, 0.2, 0.0, u0[0].cpu().numpy(), np.zeros((256, 256)), np.zeros((39, 256, 256)), False), 39 * 65536)[1],
                     "the same clip through the NumPy port of data/v2v_core_esim.py:26-69 (no noise draws)")
        c["Mpix_frames_per_s"] = c["units"] / c["seconds"] / 1e6
        res["config1_one_clip_40x256x256"]["cpu"] = c
    # the same shape batched 64x: what a DataLoader batch of small clips achieves
    fr64 = walk(64, 41, 256, 256, 2, dev)
    out64 = torch.empty((64, 8, 5, 256, 256), dtype=torch.float32, device=dev)
    p64 = torch.full((64,), 0.2, dtype=torch.float64, device=dev)
    med, mn = timeit(lambda: v2v.frames_to_voxel(fr64, p64, p64, num_bins=5, out=out64), 20)
    by = 64 * (41 * 65536 + 40 * 65536 * 4)
    res["config1_batched_64x41x256x256"] = {"ms": med, "clips_per_s": 64e3 / med, "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK}
    del fr64, out64

    # train shape: 12 clips of 201x128x128 (config/train_v2v_e2vid_10k.yaml)
    frt = walk(12, 201, 128, 128, 3, dev)
    outt = torch.empty((12, 40, 5, 128, 128), dtype=torch.float32, device=dev)
    c = lambda v, n: torch.full((n,), v, dtype=torch.float64, device=dev)
    tp, tn_, tstd, tfrac, thstd = c(0.3, 12), c(0.4, 12), c(0.05, 12), c(0.0005, 12), c(5.0, 12)
    ft = lambda: v2v.frames_to_voxel(frt, tp, tn_, num_bins=5, noise="philox", base_noise_std=tstd, hot_pixel_fraction=tfrac,
                                     hot_pixel_std=thstd, seed=7, out=outt)
    med, mn = timeit(ft, 20)
    mg = timeit_graph(ft)
    by = 12 * 128 * 128 * (201 + 200 * 4)
    res["train_batch_12x201x128x128_philox"] = {"ms": med, "ms_graph_replay": mg, "clips_per_s": 12e3 / med,
                                                "Mpix_frames_per_s": 12 * 200 * 16384 / med / 1e3,
                                                "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK,
                                                "frac_graph_replay": by / mg / 1e6 / PEAK,
                                                "note": "the real training batch; 196k pixels of parallelism only; ms = one call from "
                                                        "Python (ctypes + tensor bookkeeping between launches), ms_graph_replay = the same "
                                                        "call replayed from a CUDA graph"}
    if cpu_arm:
        vt = frt[0].cpu().numpy()
        rs = np.random.RandomState(0)

        def one_train_clip():
            u, hot, g = orc.esim_draw_randomness(201, 128, 128, 0.0005, 5.0, rs)
            orc.bin_accumulate(orc.esim_video_to_voxel(vt, 0.3, 0.4, 0.05, u, hot, g, False), 5, 1).astype(np.float32)
            return 200 * 16384
        cc = _cpu_arm(one_train_clip, "one clip [201,128,128] of the batch through the NumPy port incl. MT19937 draws")
        cc["Mpix_frames_per_s"] = cc["units"] / cc["seconds"] / 1e6
        res["train_batch_12x201x128x128_philox"]["cpu"] = cc
    del frt, outt

    # config 3: v2e core, noisy preset, HDR-degraded clips 8 x 121x480x640
    fr3 = walk(8, 121, 480, 640, 4, dev)
    fr3 = ((fr3.float() - 127.5) * 2.0 + 127.5).clamp_(0, 255).to(torch.uint8)
    g = np.random.Generator(np.random.PCG64(0))
    m = g.normal(0.2, 0.05, (8, 480, 640))
    dd = g.normal(0.0, 0.05, (8, 480, 640))
    pt = torch.from_numpy(np.clip(m + dd / 2, 0.01, None)).to(dev)
    nt = torch.from_numpy(np.clip(m - dd / 2, 0.01, None)).to(dev)
    nr = torch.from_numpy(np.exp(np.log(10) * 0.1 * g.standard_normal((8, 480, 640)).astype(np.float32)).astype(np.float32)).to(dev)
    f3 = lambda: frames_to_voxel_v2e(fr3, pt, nt, fps=24, num_bins=5, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0,
                                     leak_jitter_fraction=0.1, noise_rate=nr, noise="philox", seed=3)
    med, mn = timeit(f3, 5, 2)
    by = 8 * 307200 * (121 + 120 * 4 + 20)
    res["config3_v2e_noisy_8x121x480x640"] = {"ms": med, "clips_per_s": 8e3 / med, "Mpix_frames_per_s": 8 * 120 * 307200 / med / 1e3,
                                              "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK,
                                              "note": "includes the shot-scale pre-pass; per-pixel double-precision exp/Poisson: compute bound"}
    f3c = lambda: frames_to_voxel_v2e(fr3, pt, nt, fps=24, num_bins=5, noise="none")
    med, mn = timeit(f3c, 5, 2)
    res["config3_v2e_clean_8x121x480x640"] = {"ms": med, "clips_per_s": 8e3 / med, "Mpix_frames_per_s": 8 * 120 * 307200 / med / 1e3,
                                              "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK}
    if cpu_arm:
        v3 = fr3[0, :11].cpu().numpy()

        def v2e_sample():
            prm = dict(threshold_model="pn_related", thres_mean_mean=0.2, thres_mean_std=0.05, thres_diff_mean=0.0, thres_diff_std=0.05,
                       cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1, noise_rate_cov_decades=0.1)
            orc.v2e_video_to_voxel(v3.astype(np.float64), 24, prm, rs=np.random.RandomState(3))
            return 10 * 307200
        try:
            cc = _cpu_arm(v2e_sample, "11 frames of one HDR-degraded clip through the NumPy port of data/v2v_core_v2e.py (noisy preset)")
            cc["Mpix_frames_per_s"] = cc["units"] / cc["seconds"] / 1e6
            res["config3_v2e_noisy_8x121x480x640"]["cpu"] = cc
        except Exception as e:      # the port's entry point is optional here
            res["config3_v2e_noisy_8x121x480x640"]["cpu"] = {"unavailable": f"{type(e).__name__}: {e}"}
    del fr3, pt, nt, nr

    # config 4: 10 M events, 260x346, 400 windows
    g = np.random.Generator(np.random.PCG64(5))
    ne, h, w, wn = 10_000_000, 260, 346, 400
    xs_h, ys_h = g.integers(0, w, ne).astype(np.int16), g.integers(0, h, ne).astype(np.int16)
    ts_h, ps_h = np.sort(g.random(ne) * 10.0), (g.random(ne) < 0.5).astype(np.uint8)
    xs, ys, ts, ps = (torch.from_numpy(x).to(dev) for x in (xs_h, ys_h, ts_h, ps_h))
    off_h = np.linspace(0, ne, wn + 1).astype(np.int64)
    off = torch.from_numpy(off_h).to(dev)
    for bins in (5, 15):
        outv = torch.empty((wn, bins, h, w), dtype=torch.float32, device=dev)
        for mode in ("h5_discrete", "h5_interp"):
            med, mn = timeit(lambda: v2v.voxelize_windows(xs, ys, ts, ps, off, bins, h, w, mode=mode, out=outv, validate=False), 10)
            by = ne * (2 + 2 + 8 + 1) + wn * bins * h * w * 4
            res[f"config4_scatter_{mode}_bins{bins}"] = {"ms": med, "Mev_per_s": ne / med / 1e3, "GBps": by / med / 1e6,
                                                         "frac": by / med / 1e6 / PEAK, "algorithmic_bytes": by}
        del outv
    if cpu_arm:
        for interp in (False, True):
            def mv():
                for k in range(20):
                    a0, a1 = off_h[k], off_h[k + 1]
                    orc.make_voxel(ts_h[a0:a1], xs_h[a0:a1], ys_h[a0:a1], ps_h[a0:a1], 5, h, w, interp)
                return int(off_h[20])
            cc = _cpu_arm(mv, "the first 20 windows (500 k events) through the NumPy port of TestH5Dataset.make_voxel (np.add.at)")
            cc["Mev_per_s"] = cc["units"] / cc["seconds"] / 1e6
            res[f"config4_scatter_{'h5_interp' if interp else 'h5_discrete'}_bins5"]["cpu"] = cc
    # 1 % of events on one hot pixel
    xs2, ys2 = xs.clone(), ys.clone()
    sel = torch.from_numpy(g.random(ne) < 0.01).to(dev)
    xs2[sel], ys2[sel] = 100, 100
    outv = torch.empty((wn, 5, h, w), dtype=torch.float32, device=dev)
    med, mn = timeit(lambda: v2v.voxelize_windows(xs2, ys2, ts, ps, off, 5, h, w, mode="h5_interp", out=outv, validate=False), 10)
    res["config4_scatter_h5_interp_bins5_hotpixel"] = {"ms": med, "Mev_per_s": ne / med / 1e3}
    # legacy torch flavour, one window of 10 M events (the offline cache builder's shape)
    tsf = ts.to(torch.float32)
    pf = ps.to(torch.float32) * 2 - 1
    for bil in (True, False):
        med, mn = timeit(lambda: v2v.events_to_voxel_torch(xs, ys, tsf, pf, 5, sensor_size=(h, w), temporal_bilinear=bil), 5)
        res[f"config4_events_to_voxel_torch_bilinear{int(bil)}"] = {"ms": med, "Mev_per_s": ne / med / 1e3}
    del xs, ys, ts, ps, xs2, ys2, outv, tsf, pf

    # config 5 shape: 1080p clips, 26 frames -> 5 voxels of 5 bins, padded to /16 with fused frame output
    fr5 = walk(4, 26, 1080, 1920, 6, dev)
    vz = v2v.V2VVoxelizer(dict(num_bins=5, base_noise_std_range=[0, 0.1], hot_pixel_std_range=[0, 10]), device=dev)
    params = vz.sample_batch_params(4, rs=np.random.RandomState(0))
    store = torch.zeros((4, 5, 5, 1088, 1920), dtype=torch.float32, device=dev)
    f5 = lambda: vz.batch_to_tensors(fr5, params, seed=1, pad_multiple=16, with_stats=True, out=store)
    med, mn = timeit(f5, 10)
    by = 4 * 1080 * 1920 * (26 + 25 * 4 + 5 * 4)
    res["config5_1080p_4x26x1080x1920_padded_frames_stats"] = {"ms": med, "clips_per_s": 4e3 / med, "Mpix_frames_per_s": 4 * 25 * 1080 * 1920 / med / 1e3,
                                                               "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK,
                                                               "note": "voxels written straight into the /16-padded consumer layout + frame/255 output"}
    if cpu_arm:
        v5 = fr5[0, :6].cpu().numpy()
        rs5 = np.random.RandomState(1)
        p5 = params[0]

        def c5():
            u, hot, gg = orc.esim_draw_randomness(6, 1080, 1920, p5["hot_pixel_fraction"], p5["hot_pixel_std"], rs5)
            orc.bin_accumulate(orc.esim_video_to_voxel(v5, p5["pos_thres"], p5["neg_thres"], p5["base_noise_std"], u, hot, gg, False), 5, 1).astype(np.float32)
            return 5 * 1080 * 1920
        cc = _cpu_arm(c5, "6 frames (one voxel) of one 1080p clip through the NumPy port incl. MT19937 draws")
        cc["Mpix_frames_per_s"] = cc["units"] / cc["seconds"] / 1e6
        res["config5_1080p_4x26x1080x1920_padded_frames_stats"]["cpu"] = cc
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    res = secondary_configs(torch.device("cuda:0"), cpu_arm=not a.no_cpu)
    print(json.dumps(res, indent=1))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
