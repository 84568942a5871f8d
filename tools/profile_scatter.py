#!/usr/bin/env python
"""ncu driver: one config-4 scatter launch per mode (10 M events, 260x346, 400 windows)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import v2v_b200 as v2v
dev = torch.device("cuda:0")
g = np.random.Generator(np.random.PCG64(5))
ne, h, w, wn = 10_000_000, 260, 346, 400
xs = torch.from_numpy(g.integers(0, w, ne).astype(np.int16)).to(dev)
ys = torch.from_numpy(g.integers(0, h, ne).astype(np.int16)).to(dev)
ts = torch.from_numpy(np.sort(g.random(ne) * 10.0)).to(dev)
ps = torch.from_numpy((g.random(ne) < 0.5).astype(np.uint8)).to(dev)
off = torch.from_numpy(np.linspace(0, ne, wn + 1).astype(np.int64)).to(dev)
bins = int(sys.argv[1]) if len(sys.argv) > 1 else 5
out = torch.empty((wn, bins, h, w), dtype=torch.float32, device=dev)
for mode in ("h5_discrete", "h5_interp"):
    for _ in range(2):
        v2v.voxelize_windows(xs, ys, ts, ps, off, bins, h, w, mode=mode, out=out, validate=False)
torch.cuda.synchronize()
