#!/usr/bin/env python
"""ncu / timing driver for the small-launch regime: the shipped training batch (12 clips of 201x128x128) and BASELINE
config 1 (one clip of 40x256x256).  `--time` prints CUDA-event timings for every kernel choice."""
import os, sys, argparse
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import v2v_b200 as v2v
from v2v_b200 import _lib
from tools.bench_configs import walk, timeit
ap = argparse.ArgumentParser()
ap.add_argument("--time", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
frt = walk(12, 201, 128, 128, 3, dev)
outt = torch.empty((12, 40, 5, 128, 128), dtype=torch.float32, device=dev)
c = lambda v, n: torch.full((n,), v, dtype=torch.float64, device=dev)
kw = dict(num_bins=5, noise="philox", base_noise_std=c(0.05, 12), hot_pixel_fraction=c(0.0005, 12), hot_pixel_std=c(5.0, 12), out=outt)
f = lambda flags=0: v2v.frames_to_voxel(frt, c(0.3, 12), c(0.4, 12), kernel_flags=flags, **kw)
by = 12 * 128 * 128 * (201 + 200 * 4)
if a.time:
    for name, fl in (("default", 0), ("p1", _lib.ESIM_FLAG_SMALL_P1), ("generic", _lib.ESIM_FLAG_GENERIC)) + tuple((f"geom{g}", _lib.esim_flag_geom(g)) for g in range(1, 8)):
        try:
            med, mn = timeit(lambda: f(fl), 20)
            print(f"train batch {name:8s} ms={med:.4f} GB/s={by / med / 1e6:.0f}")
        except Exception as e:
            print(name, "failed", e)
else:
    for _ in range(3):
        f()
    torch.cuda.synchronize()
