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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    res = {"peak_gbs": PEAK, "gpu": torch.cuda.get_device_name(0)}

    # config 1: one 40x256x256 clip, raw core (num_bins=1), fixed thresholds, no noise: latency bound
    fr = walk(1, 40, 256, 256, 1, dev)
    u0 = torch.rand((1, 256, 256), dtype=torch.float64, device=dev)
    out1 = torch.empty((1, 39, 1, 256, 256), dtype=torch.float32, device=dev)
    pos = torch.full((1,), 0.2, dtype=torch.float64, device=dev)
    med, mn = timeit(lambda: v2v.frames_to_voxel(fr, pos, pos, num_bins=1, u0=u0, out=out1), 20)
    by = 40 * 65536 + 39 * 65536 * 4
    res["config1_one_clip_40x256x256"] = {"ms": med, "ms_min": mn, "clips_per_s": 1e3 / med, "Mpix_frames_per_s": 39 * 65536 / med / 1e3,
                                          "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK, "note": "single small clip: launch/latency bound (SURVEY §7)"}
    # the same shape batched 64x: what a DataLoader batch of small clips achieves
    fr64 = walk(64, 41, 256, 256, 2, dev)
    out64 = torch.empty((64, 8, 5, 256, 256), dtype=torch.float32, device=dev)
    p64 = torch.full((64,), 0.2, dtype=torch.float64, device=dev)
    med, mn = timeit(lambda: v2v.frames_to_voxel(fr64, p64, p64, num_bins=5, out=out64), 20)
    by = 64 * (41 * 65536 + 40 * 65536 * 4)
    res["config1_batched_64x41x256x256"] = {"ms": med, "clips_per_s": 64e3 / med, "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK}

    # train shape: 12 clips of 201x128x128 (config/train_v2v_e2vid_10k.yaml)
    frt = walk(12, 201, 128, 128, 3, dev)
    outt = torch.empty((12, 40, 5, 128, 128), dtype=torch.float32, device=dev)
    c = lambda v, n: torch.full((n,), v, dtype=torch.float64, device=dev)
    med, mn = timeit(lambda: v2v.frames_to_voxel(frt, c(0.3, 12), c(0.4, 12), num_bins=5, noise="philox", base_noise_std=c(0.05, 12),
                                                 hot_pixel_fraction=c(0.0005, 12), hot_pixel_std=c(5.0, 12), out=outt), 20)
    by = 12 * 128 * 128 * (201 + 200 * 4)
    res["train_batch_12x201x128x128_philox"] = {"ms": med, "clips_per_s": 12e3 / med, "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK,
                                                "note": "the real training batch; 196k pixels of parallelism only"}

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
    res["config3_v2e_clean_8x121x480x640"] = {"ms": med, "clips_per_s": 8e3 / med, "GBps": by / med / 1e6, "frac": by / med / 1e6 / PEAK}

    # config 4: 10 M events, 260x346, 400 windows
    g = np.random.Generator(np.random.PCG64(5))
    ne, h, w, wn = 10_000_000, 260, 346, 400
    xs = torch.from_numpy(g.integers(0, w, ne).astype(np.int16)).to(dev)
    ys = torch.from_numpy(g.integers(0, h, ne).astype(np.int16)).to(dev)
    ts = torch.from_numpy(np.sort(g.random(ne) * 10.0)).to(dev)
    ps = torch.from_numpy((g.random(ne) < 0.5).astype(np.uint8)).to(dev)
    off = torch.from_numpy(np.linspace(0, ne, wn + 1).astype(np.int64)).to(dev)
    for bins in (5, 15):
        outv = torch.empty((wn, bins, h, w), dtype=torch.float32, device=dev)
        for mode in ("h5_discrete", "h5_interp"):
            med, mn = timeit(lambda: v2v.voxelize_windows(xs, ys, ts, ps, off, bins, h, w, mode=mode, out=outv), 10)
            by = ne * (2 + 2 + 8 + 1) + wn * bins * h * w * 4
            res[f"config4_scatter_{mode}_bins{bins}"] = {"ms": med, "Mev_per_s": ne / med / 1e3, "GBps": by / med / 1e6,
                                                         "frac": by / med / 1e6 / PEAK, "algorithmic_bytes": by}
    # 1 % of events on one hot pixel
    xs2, ys2 = xs.clone(), ys.clone()
    sel = torch.from_numpy(g.random(ne) < 0.01).to(dev)
    xs2[sel], ys2[sel] = 100, 100
    outv = torch.empty((wn, 5, h, w), dtype=torch.float32, device=dev)
    med, mn = timeit(lambda: v2v.voxelize_windows(xs2, ys2, ts, ps, off, 5, h, w, mode="h5_interp", out=outv), 10)
    res["config4_scatter_h5_interp_bins5_hotpixel"] = {"ms": med, "Mev_per_s": ne / med / 1e3}
    # legacy torch flavour, one window of 10 M events (the offline cache builder's shape)
    tsf = ts.to(torch.float32)
    pf = ps.to(torch.float32) * 2 - 1
    for bil in (True, False):
        med, mn = timeit(lambda: v2v.events_to_voxel_torch(xs, ys, tsf, pf, 5, sensor_size=(h, w), temporal_bilinear=bil), 5)
        res[f"config4_events_to_voxel_torch_bilinear{int(bil)}"] = {"ms": med, "Mev_per_s": ne / med / 1e3}

    # config 5 shape: 1080p clips, 26 frames -> 5 voxels of 5 bins, padded to /16 with fused frame output
    fr5 = walk(4, 26, 1080, 1920, 6, dev)
    vz = v2v.V2VVoxelizer(dict(num_bins=5, base_noise_std_range=[0, 0.1], hot_pixel_std_range=[0, 10]), device=dev)
    params = vz.sample_batch_params(4, rs=np.random.RandomState(0))
    store = torch.zeros((4, 5, 5, 1088, 1920), dtype=torch.float32, device=dev)
    f5 = lambda: vz.batch_to_tensors(fr5, params, seed=1, pad_multiple=16, with_stats=True, out=store)
    med, mn = timeit(f5, 10)
    by = 4 * 1080 * 1920 * (26 + 25 * 4 + 5 * 4)
    res["config5_1080p_4x26x1080x1920_padded_frames_stats"] = {"ms": med, "clips_per_s": 4e3 / med, "GBps": by / med / 1e6,
                                                               "frac": by / med / 1e6 / PEAK,
                                                               "note": "voxels written straight into the /16-padded consumer layout + frame/255 output"}
    print(json.dumps(res, indent=1))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
