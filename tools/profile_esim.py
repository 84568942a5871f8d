#!/usr/bin/env python
"""Small driver for ncu: launches the ESIM kernel a few times on BASELINE config-2 clips.

    ncu --set full ... python tools/profile_esim.py --noise philox --clips 8 --iters 3
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import v2v_b200 as v2v  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--noise", default="philox")
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--time", action="store_true")
ap.add_argument("--stats", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
frames = bench.make_clips(torch, dev, a.clips, 100)
vz = v2v.V2VVoxelizer(bench.TRAIN_CFG, device=dev)
params = vz.sample_batch_params(a.clips, rs=np.random.RandomState(1234))
col = lambda k: torch.tensor([p[k] for p in params], dtype=torch.float64, device=dev)
out = torch.empty((a.clips, 24, 5, bench.H, bench.W), dtype=torch.float32, device=dev)
kw = dict(num_bins=5, out=out, with_stats=a.stats)
if a.noise == "philox":
    kw.update(noise="philox", base_noise_std=col("base_noise_std"), hot_pixel_fraction=col("hot_pixel_fraction"),
              hot_pixel_std=col("hot_pixel_std"), seed=1)
evs = []
for i in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v2v.frames_to_voxel(frames, col("pos_thres"), col("neg_thres"), **kw)
    e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
if a.time:
    ms = [x.elapsed_time(y) for x, y in evs]
    gb = a.clips * bench.ALGO_BYTES_PER_CLIP / 1e9
    print(f"noise={a.noise} stats={int(a.stats)} clips={a.clips} geom={os.environ.get('V2V_ESIM_GEOM','default')} staged={os.environ.get('V2V_ESIM_STAGED','0')} "
          f"ms={np.min(ms[1:] or ms):.3f} GB/s={gb / (np.min(ms[1:] or ms) * 1e-3):.0f}")
