#!/usr/bin/env python
"""Same-box A/B of library builds on the config-5 shape (4 x [26,1080,1920], padded, frames out, statistics, Philox) and the
training batch (12 x [201,128,128]): python tools/ab_c5.py libA.so libB.so ...  (each in a fresh subprocess, 3 rounds)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
import v2v_b200 as v2v
from tools.bench_configs import walk, timeit_graph
dev = torch.device("cuda:0")
c = lambda v, n: torch.full((n,), v, dtype=torch.float64, device=dev)
fr = walk(4, 26, 1080, 1920, 5, dev)
p, n, s, f, h = c(0.3, 4), c(0.4, 4), c(0.05, 4), c(0.0005, 4), c(5.0, 4)
f5 = lambda: v2v.frames_to_voxel(fr, p, n, num_bins=5, noise="philox", base_noise_std=s, hot_pixel_fraction=f, hot_pixel_std=h, seed=3,
                                 pad_multiple=16, frame_out="frames", with_stats=True)
f5()
ms5 = timeit_graph(f5, 10)
frt = walk(12, 201, 128, 128, 3, dev)
outt = torch.empty((12, 40, 5, 128, 128), dtype=torch.float32, device=dev)
tp, tn, ts, tf, th = c(0.3, 12), c(0.4, 12), c(0.05, 12), c(0.0005, 12), c(5.0, 12)
ft = lambda: v2v.frames_to_voxel(frt, tp, tn, num_bins=5, noise="philox", base_noise_std=ts, hot_pixel_fraction=tf, hot_pixel_std=th, seed=7, out=outt)
mst = timeit_graph(ft, 20)
print("c5_ms=%%.4f train_ms=%%.4f" %% (ms5, mst))
''' % ROOT
libs = sys.argv[1:]
res = {l: [] for l in libs}
for rnd in range(3):
    for l in libs:
        out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=dict(os.environ, V2V_B200_LIB=os.path.abspath(l)))
        res[l].append(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:])
for l in libs:
    print(l, res[l])
