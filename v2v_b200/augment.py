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


_calls = 0      # per-process call counter: the default noise stream of the philox mode (never repeats a field)


def _fresh_seed():
    """A seed for the in-kernel generator when the caller gives none: drawn from ``np.random`` like every draw of the
    reference's function, so a DataLoader worker's seeding governs it and no two calls share a noise field."""
    return int(np.random.randint(0, 2 ** 31 - 1)) << 32 | int(np.random.randint(0, 2 ** 31 - 1))


def add_noise_to_voxel(voxel: torch.Tensor, noise_std=1.0, noise_fraction=0.1, integer_noise=False, *, rng="philox",
                       seed=None, stream_id=None) -> torch.Tensor:
    """In place on a CUDA float32 voxel.  rng="philox": noise generated in the kernel under ``seed`` / ``stream_id``;
    with ``seed=None`` (the reference's signature has no seed) a fresh seed is drawn from ``np.random`` on every call and
    ``stream_id`` defaults to a per-process call counter, so default calls never repeat a noise field."""
    global _calls
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
        seed = stream_id = 0
    elif rng == "philox":
        if seed is None:
            seed = _fresh_seed()
        if stream_id is None:
            stream_id = _calls
        _calls += 1
    else:
        raise ValueError("rng must be 'numpy' or 'philox'")
    s = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_voxel_add_noise(_p(voxel), n, _p(noise_t), _p(mask_t), float(noise_std), float(noise_fraction),
                                                   int(bool(integer_noise)), int(rng == "philox"), int(seed) & (2 ** 64 - 1),
                                                   int(stream_id), C.c_void_p(s.cuda_stream)))
    return voxel


def add_hot_pixels_to_voxels(voxels: torch.Tensor, hot_pixel_std=1.0, max_hot_pixel_fraction=0.001, integer_noise=False) -> torch.Tensor:
    """voxels ``[T,C,H,W]`` CUDA float32, in place.  The handful of hot-pixel draws use the reference's host generators
    in the reference's order (``random.uniform``, ``np.random.randint`` x2, then the values, :10-24) — same seeds, same
    result.  With ``integer_noise`` the reference overwrites its row coordinates with the Poisson draws (:19), so the
    noise lands on rows 0, 1, 2, ...; that is what runs in the reference's training, and it is reproduced here."""
    if not voxels.is_cuda or voxels.dtype != torch.float32 or not voxels.is_contiguous() or voxels.dim() != 4:
        raise _lib.V2VError(-1, "voxels must be a contiguous CUDA float32 [T,C,H,W] tensor")
    T, Cc, H, W = voxels.shape
    frac = random.uniform(0, max_hot_pixel_fraction)
    num = int(frac * H * W)
    x = np.random.randint(0, W, num)
    y = np.random.randint(0, H, num)
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * hot_pixel_std ** 2)) / 2
        y = np.random.poisson(lam=lmb, size=num)                               # sic: the rows become the Poisson draws
        sign = 2 * np.random.randint(0, 2, size=num) - 1
        val = (y * sign).astype(np.float64)
        if num and int(y.max()) >= H:
            raise IndexError(f"index {int(y.max())} is out of bounds for axis 0 with size {H}")   # np.add.at in the reference
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


def cached_sequence_item(all_frame: torch.Tensor, all_flow: torch.Tensor, all_voxel: torch.Tensor, sequence_length: int, *,
                         proba_pause_when_running=0.05, proba_pause_when_paused=0.9, noise_std=0.1, noise_fraction=1.0,
                         hot_pixel_std=0.1, max_hot_pixel_fraction=0.001, integer_noise=False, rng="philox", seed=None):
    """The augmentation of ``ESIMH5Dataset.__getitem__`` (data/esim_dataset.py:108-143) on the GPU, after the crop and
    flip (which are views): pause sequence, per-step voxel noise, hot pixels.  Inputs are CUDA float32
    ``[S,1,H,W] / [S,2,H,W] / [S,C,H,W]`` slices of the cached-voxel file (``frames / flow / events`` datasets written by
    scripts/esim_to_voxel.py:45-51), S >= sequence_length.  Returns ``{"frame","flow","events"}`` of ``sequence_length``
    steps, float32 — the reference's item without ``data_source_idx``.

    The pause decisions are one ``np.random.rand()`` per step in the reference's order.  rng="numpy": the voxel noise of
    every step is drawn on the host right after that step's pause draw, exactly as the reference interleaves them
    (same seeds, same item; host bound).  rng="philox": the noise of the whole sequence is generated in the kernel."""
    for t in (all_frame, all_flow, all_voxel):
        if not t.is_cuda or t.dtype != torch.float32:
            raise _lib.V2VError(-1, "all_frame / all_flow / all_voxel must be CUDA float32 tensors (no CPU fallback)")
    dev = all_voxel.device
    T = int(sequence_length)
    vshape = tuple(all_voxel.shape[1:])
    src = np.full(T, -1, dtype=np.int64)
    fsrc = np.zeros(T, dtype=np.int64)
    noise = np.empty((T,) + vshape, dtype=np.float64) if rng == "numpy" else None
    mask = np.empty((T,) + vshape, dtype=np.float64) if (rng == "numpy" and noise_fraction < 1.0) else None
    paused, k = False, 0
    for t in range(T):
        u = np.random.rand()                                                   # :116
        paused = u < (proba_pause_when_paused if paused else proba_pause_when_running)
        if t > 0 and paused:                                                   # :122-126: repeat the frame, zero flow / voxel
            fsrc[t] = fsrc[t - 1]
        else:
            src[t] = fsrc[t] = k
            k += 1
        if rng == "numpy":                                                     # :136, the draws of add_noise_to_voxel
            if integer_noise:
                lmb = (-1 + np.sqrt(1 + 4 * noise_std ** 2)) / 2
                y = np.random.poisson(lam=lmb, size=vshape)
                noise[t] = y * (2 * np.random.randint(0, 2, size=vshape) - 1)
            else:
                noise[t] = noise_std * np.random.randn(*vshape)
            if mask is not None:
                mask[t] = np.random.rand(*vshape)
    src_t = torch.from_numpy(src).to(dev)
    keep = (src_t >= 0)
    idx = src_t.clamp(min=0)
    frame = all_frame.index_select(0, torch.from_numpy(fsrc).to(dev)).contiguous()
    flow = (all_flow.index_select(0, idx) * keep.view(-1, 1, 1, 1).to(all_flow.dtype)).contiguous()
    voxel = (all_voxel.index_select(0, idx) * keep.view(-1, 1, 1, 1).to(all_voxel.dtype)).contiguous()
    s = torch.cuda.current_stream(dev)
    if rng == "numpy":
        noise_t = torch.from_numpy(noise).to(dev)
        mask_t = torch.from_numpy(mask).to(dev) if mask is not None else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().v2v_voxel_add_noise(_p(voxel), voxel.numel(), _p(noise_t), _p(mask_t), float(noise_std),
                                                       float(noise_fraction), int(bool(integer_noise)), 0, 0, 0,
                                                       C.c_void_p(s.cuda_stream)))
    else:
        add_noise_to_voxel(voxel, noise_std, noise_fraction, integer_noise, rng="philox", seed=seed)
    add_hot_pixels_to_voxels(voxel, hot_pixel_std, max_hot_pixel_fraction, integer_noise)      # :138
    return {"frame": frame, "flow": flow, "events": voxel}
