"""Time / profile the v2e kernel alone (BASELINE config 3 shape).

    python tools/profile_v2e.py --preset noisy --clips 8 --time
    ncu --set full -k regex:v2e_kernel -c 1 ... python tools/profile_v2e.py --preset noisy --iters 1
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_configs import walk  # noqa: E402
from v2v_b200.v2e import frames_to_voxel_v2e  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="noisy", choices=["noisy", "clean", "cutoff", "leak", "shot"])
    ap.add_argument("--clips", type=int, default=8)
    ap.add_argument("--frames", type=int, default=121)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--stats", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, N, H, W = a.clips, a.frames, 480, 640
    fr = walk(B, N, H, W, 4, dev)
    fr = ((fr.float() - 127.5) * 2.0 + 127.5).clamp_(0, 255).to(torch.uint8)       # HDR degrade, data/v2v_datasets.py:473-477
    g = np.random.Generator(np.random.PCG64(0))
    m = g.normal(0.2, 0.05, (B, H, W))
    dd = g.normal(0.0, 0.05, (B, H, W))
    pt = torch.from_numpy(np.clip(m + dd / 2, 0.01, None)).to(dev)
    nt = torch.from_numpy(np.clip(m - dd / 2, 0.01, None)).to(dev)
    nr = torch.from_numpy(np.exp(np.log(10) * 0.1 * g.standard_normal((B, H, W)).astype(np.float32)).astype(np.float32)).to(dev)
    kw = {
        "noisy": dict(cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1, noise_rate=nr, noise="philox", seed=3),
        "clean": dict(noise="none"),
        "cutoff": dict(cutoff_hz=30.0, noise="none"),
        "leak": dict(leak_rate_hz=0.1, leak_jitter_fraction=0.1, noise_rate=nr, noise="philox", seed=3),
        "shot": dict(shot_noise_rate_hz=5.0, noise="philox", seed=3),
    }[a.preset]
    fn = lambda: frames_to_voxel_v2e(fr, pt, nt, fps=24, num_bins=5, with_stats=a.stats, **kw)  # noqa: E731
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    if a.time:
        med = float(np.median(ms))
        by = B * H * W * (N + (N - 1) * 4 + 20)
        print(f"v2e preset={a.preset} clips={B} ms={med:.3f} clips/s={B * 1e3 / med:.0f} GB/s={by / med / 1e6:.0f} "
              f"sum={float(out['voxel'].double().abs().sum()):.0f}")


if __name__ == "__main__":
    main()
