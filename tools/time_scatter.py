#!/usr/bin/env python
"""Config-4 scatter timings (10 M events, 260x346, 400 windows), CUDA events, every mode; `--check` compares the interpolated
mode with the contiguous-range kernel."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import v2v_b200 as v2v
from v2v_b200 import _lib
dev = torch.device("cuda:0")
g = np.random.Generator(np.random.PCG64(5))
ne, h, w, wn = 10_000_000, 260, 346, 400
xs = torch.from_numpy(g.integers(0, w, ne).astype(np.int16)).to(dev)
ys = torch.from_numpy(g.integers(0, h, ne).astype(np.int16)).to(dev)
ts = torch.from_numpy(np.sort(g.random(ne) * 10.0)).to(dev)
ps = torch.from_numpy((g.random(ne) < 0.5).astype(np.uint8)).to(dev)
off = torch.from_numpy(np.linspace(0, ne, wn + 1).astype(np.int64)).to(dev)
peak = 6557.4
for bins in (5, 15):
    out = torch.empty((wn, bins, h, w), dtype=torch.float32, device=dev)
    algo = ne * 13 + out.numel() * 4
    for mode in ("h5_discrete", "h5_interp"):
        for flags, tag in ((0, "default"), (_lib.SCATTER_FLAG_RANGES, "ranges")):
            if mode == "h5_discrete" and flags:
                continue
            ms = []
            for i in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                v2v.voxelize_windows(xs, ys, ts, ps, off, bins, h, w, mode=mode, out=out, validate=False, kernel_flags=flags)
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            m = min(ms[2:])
            print(f"bins={bins} {mode} {tag}: {m:.3f} ms  {ne / m / 1e6:.1f} Gev/s  frac={algo / (m * 1e-3) / 1e9 / peak:.3f}")
if "--check" in sys.argv:
    a = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_interp", validate=False)
    b = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_interp", validate=False, kernel_flags=_lib.SCATTER_FLAG_RANGES)
    print("max |sorted - ranges| =", float((a - b).abs().max()))
