"""Voxel-space noise augmentation of cached voxels on the GPU.

Same names and arguments as the reference's ``data/esim_dataset.py``: ``add_noise_to_voxel`` (:33-46) and
``add_hot_pixels_to_voxels`` (:7-30), operating in place on CUDA float32 voxels.

rng="numpy": every random field is drawn on the host from the same generators, in the same order, as the reference
(``np.random`` and, for the hot-pixel fraction, Python's ``random``) and applied on the GPU — same seeds, same result
(float64 add, float32 store).  rng="philox": the per-element noise is generated in the kernel (throughput mode); the
few hot-pixel draws stay on the host.
"""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import _lib
from .events import _image


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def add_noise_to_voxel(voxel: torch.Tensor, noise_std=1.0, noise_fraction=0.1, integer_noise=False, *, rng="philox",
                       seed: int = 0, stream_id: int = 0) -> torch.Tensor:
    if not voxel.is_cuda or voxel.dtype != torch.float32 or not voxel.is_contiguous():
        raise _lib.V2VError(-1, "voxel must be a contiguous CUDA float32 tensor (no CPU fallback)")
    dev = voxel.device
    n = voxel.numel()
    noise_t = mask_t = None
    if rng == "numpy":
        shape = tuple(voxel.shape)
        if integer_noise:                                                      # :35-39
            lmb = (-1 + np.sqrt(1 + 4 * noise_std ** 2)) / 2
            y = np.random.poisson(lam=lmb, size=shape)
            sign = 2 * np.random.randint(0, 2, size=shape) - 1
            noise = (y * sign).astype(np.float64)
        else:
            noise = noise_std * np.random.randn(*shape)                        # :41
        noise_t = torch.from_numpy(np.ascontiguousarray(noise)).to(dev)
        if noise_fraction < 1.0:
            mask_t = torch.from_numpy(np.random.rand(*shape)).to(dev)          # :44
    elif rng != "philox":
        raise ValueError("rng must be 'numpy' or 'philox'")
    s = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_voxel_add_noise(_p(voxel), n, _p(noise_t), _p(mask_t), float(noise_std), float(noise_fraction),
                                                   int(bool(integer_noise)), int(rng == "philox"), int(seed) & (2 ** 64 - 1),
                                                   int(stream_id), C.c_void_p(s.cuda_stream)))
    return voxel


def add_hot_pixels_to_voxels(voxels: torch.Tensor, hot_pixel_std=1.0, max_hot_pixel_fraction=0.001, integer_noise=False) -> torch.Tensor:
    """voxels ``[T,C,H,W]`` CUDA float32, in place.  The handful of hot-pixel draws use the reference's host generators
    in the reference's order (``random.uniform``, ``np.random.randint`` x2, then the values, :10-24)."""
    if not voxels.is_cuda or voxels.dtype != torch.float32 or not voxels.is_contiguous() or voxels.dim() != 4:
        raise _lib.V2VError(-1, "voxels must be a contiguous CUDA float32 [T,C,H,W] tensor")
    T, Cc, H, W = voxels.shape
    frac = random.uniform(0, max_hot_pixel_fraction)
    num = int(frac * H * W)
    x = np.random.randint(0, W, num)
    y = np.random.randint(0, H, num)
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * hot_pixel_std ** 2)) / 2
        yy = np.random.poisson(lam=lmb, size=num)
        sign = 2 * np.random.randint(0, 2, size=num) - 1
        val = (yy * sign).astype(np.float64)
    else:
        val = np.random.randn(num)
        val *= hot_pixel_std
    if num == 0:
        return voxels
    # np.add.at(noise, (y, x), val): an event image with float64 weights
    noise_map = _image(x.astype(np.int64), y.astype(np.int64), val, (H, W), False, False, False, torch.float64, voxels.device)
    s = torch.cuda.current_stream(voxels.device)
    with torch.cuda.device(voxels.device):
        _lib.check(_lib.load().v2v_voxel_add_map(_p(voxels), T * Cc, H * W, _p(noise_map), C.c_void_p(s.cuda_stream)))
    return voxels
