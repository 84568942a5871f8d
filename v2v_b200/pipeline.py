"""Host-buffer pipeline: pinned uint8 clips in, float32 voxels back in pinned host memory.

This is the end-to-end form of the drop-in (the reference's dataset returns host
tensors, data/v2v_datasets.py:351-356): the batch is cut into chunks and the
H2D copy of chunk i+1, the kernel of chunk i and the D2H copy of chunk i-1 run
concurrently on three CUDA streams over a small ring of device staging buffers.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .esim import frames_to_voxel


class HostPipeline:
    def __init__(self, voxelizer, device="cuda", clips_per_chunk: int = 1, seed: int = 0, depth: int = 3):
        self.vz = voxelizer
        self.dev = torch.device(device)
        self.chunk = int(clips_per_chunk)
        self.depth = int(depth)
        self.seed = seed
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_k = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._shape = None
        self._slots = None

    def _alloc(self, n, h, w):
        bins, fpb = self.vz.num_bins, self.vz.frames_per_bin
        t = (n - 1) // (bins * fpb)
        if self._shape != (n, h, w):
            self._slots = [dict(fin=torch.empty((self.chunk, n, h, w), dtype=torch.uint8, device=self.dev),
                                fout=torch.empty((self.chunk, t, bins, h, w), dtype=torch.float32, device=self.dev),
                                done=None) for _ in range(self.depth)]
            self._shape = (n, h, w)
        return t

    def run(self, host_frames: torch.Tensor, params: Sequence[dict], host_out: torch.Tensor,
            clip_index_base: int = 0, stats: Optional[torch.Tensor] = None):
        """host_frames: pinned uint8 [B,N,H,W]; host_out: pinned float32 [B,T,bins,H,W] (filled on return
        of ``torch.cuda.synchronize`` / ``self.synchronize()``), or None to keep the voxels on the
        device (read them from ``self.device_voxels`` chunk by chunk via ``on_chunk``).  Returns the
        per-clip stats tensor."""
        B, n, h, w = host_frames.shape
        self._alloc(n, h, w)
        bins, fpb = self.vz.num_bins, self.vz.frames_per_bin
        keys = ("pos_thres", "neg_thres", "base_noise_std", "hot_pixel_fraction", "hot_pixel_std")
        pm = torch.from_numpy(np.array([[p[k] for p in params] for k in keys], dtype=np.float64)).to(self.dev)
        all_stats = torch.zeros((B, 2), dtype=torch.int64, device=self.dev)
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_k):
            s.wait_stream(cur)
        for ci, b0 in enumerate(range(0, B, self.chunk)):
            b1 = min(B, b0 + self.chunk)
            nb = b1 - b0
            slot = self._slots[ci % self.depth]
            if slot["done"] is not None:
                self.s_in.wait_event(slot["done"])                 # staging buffers free again
                self.s_k.wait_event(slot["done"])
            with torch.cuda.stream(self.s_in):
                slot["fin"][:nb].copy_(host_frames[b0:b1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_k):
                self.s_k.wait_event(ev_in)
                o = frames_to_voxel(slot["fin"][:nb], pm[0, b0:b1], pm[1, b0:b1], num_bins=bins, frames_per_bin=fpb,
                                    noise="philox", base_noise_std=pm[2, b0:b1], hot_pixel_fraction=pm[3, b0:b1],
                                    hot_pixel_std=pm[4, b0:b1], put_noise_external=self.vz.put_noise_external,
                                    seed=self.seed, clip_index_base=clip_index_base + b0, with_stats=True,
                                    out=slot["fout"][:nb], stream=self.s_k)
                all_stats[b0:b1] = o.stats
                ev_k = torch.cuda.Event()
                ev_k.record(self.s_k)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_k)
                if host_out is not None:
                    host_out[b0:b1].copy_(slot["fout"][:nb], non_blocking=True)
                slot["done"] = torch.cuda.Event()
                slot["done"].record(self.s_out)
        # The current stream waits for the kernels (the returned statistics, ``device_voxels``) but NOT for the D2H copies:
        # host memory is only safe to read after ``synchronize()`` anyway, and without that wait the next call's H2D copies
        # and kernels run under this call's draining D2H stream (back-to-back calls keep the link busy: the bottleneck).
        cur.wait_stream(self.s_k)
        return all_stats

    def synchronize(self):
        """Block until every voxel of the calls so far is in ``host_out``."""
        self.s_out.synchronize()
