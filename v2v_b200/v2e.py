"""v2e-style video -> voxel on B200: host side.

``video_to_voxel`` keeps the reference's signature (data/v2v_core_v2e.py:556-581).
``frames_to_voxel_v2e`` is the batched GPU-resident entry.  Per-pixel threshold
maps, the noise-rate map and (in parity mode) the per-frame random fields are
prepared on the host with the reference's NumPy expressions and draw order; the
per-pixel recurrence runs in one fused kernel (csrc/v2e.cu).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

import os

from . import _lib
from .esim import _as_dev, _ptr


def _env_kernel_flags() -> int:
    """Kernel-selection knobs of the tests and sweep tools (read in the Python host, per call; the C library reads no
    environment): V2V_V2E_GENERIC=1, V2V_V2E_FAST=1, V2V_V2E_BF=0."""
    f = 0
    if os.environ.get("V2V_V2E_GENERIC"):
        f |= _lib.V2E_FLAG_GENERIC
    if os.environ.get("V2V_V2E_FAST"):
        f |= _lib.V2E_FLAG_FAST
    if os.environ.get("V2V_V2E_BF", "1") == "0":
        f |= _lib.V2E_FLAG_DIVERGENT_DIV
    return f

_NOISE = {"none": _lib.NOISE_NONE, "explicit": _lib.NOISE_EXPLICIT, "philox": _lib.NOISE_PHILOX}
TIME_INVARIANT_MODELS = ("pn_related", "spatial_independent")
PER_FRAME_MODEL = "spatial_temporal_independent"


def v2e_log_lut() -> np.ndarray:
    """float32[256] = float32(log(v/255 + 0.01)): the effective ``lin_log``
    (data/v2v_core_v2e.py:120-137; the lin/log blend is overwritten at :135)."""
    v = np.arange(256, dtype=np.float64)
    return np.log(v / 255 + 0.01).astype(np.float32)


def threshold_maps(threshold_model, thres_mean_mean, thres_mean_std, thres_diff_mean, thres_diff_std, shape,
                   rs=np.random):
    """Per-pixel (pos, neg) float64 maps, drawn and clipped like ``_init``
    (data/v2v_core_v2e.py:333-343, 392-394)."""
    if threshold_model == "pn_related":
        m = rs.normal(loc=thres_mean_mean, scale=thres_mean_std, size=shape)
        dd = rs.normal(loc=thres_diff_mean, scale=thres_diff_std, size=shape)
        pos, neg = m + (dd / 2), m - (dd / 2)
    elif threshold_model in ("spatial_independent", PER_FRAME_MODEL):
        pos = rs.normal(loc=thres_mean_mean, scale=thres_mean_std, size=shape)
        neg = rs.normal(loc=thres_mean_mean, scale=thres_mean_std, size=shape)
    else:
        raise NotImplementedError(
            f"threshold_model {threshold_model!r} is not implemented (spatial_independent_temporal_changing adds a random "
            "walk to the maps every frame; no caller in the reference, see DESIGN.md)")
    return np.clip(pos, a_min=0.01, a_max=None), np.clip(neg, a_min=0.01, a_max=None)


def noise_rate_map(noise_rate_cov_decades, shape, rs=np.random) -> np.ndarray:
    """float32 log-normal leak-rate multipliers (data/v2v_core_v2e.py:348-349)."""
    nr = rs.randn(*shape).astype(np.float32)
    return np.exp(math.log(10) * noise_rate_cov_decades * nr)


def frames_to_voxel_v2e(frames: torch.Tensor, pos_thres, neg_thres, *, fps: float, num_bins: int = 1,
                        frames_per_bin: int = 1, cutoff_hz: float = 0.0, leak_rate_hz: float = 0.0,
                        shot_noise_rate_hz: float = 0.0, leak_jitter_fraction: float = 0.0, noise_rate=None,
                        pos_thres_nominal: float = 0.2, neg_thres_nominal: float = 0.2, noise: str = "none",
                        leak_randn=None, pos_shot=None, neg_shot=None, seed: int = 0, clip_index_base: int = 0,
                        with_stats: bool = False, lut: Optional[np.ndarray] = None,
                        return_fields: bool = False, frame_index=None, value_map=None,
                        u8_intensity: bool = False, kernel_flags: int = 0) -> dict:
    """CUDA uint8 ``[B,N,H,W]`` + per-pixel threshold maps ``[B,H,W]`` -> float32 ``[B,T,bins,H,W]``.
    ``u8_intensity``: the reference was handed a uint8 video, so its ``rescale_intensity_frame``
    (data/v2v_core_v2e.py:190) wrapped ``new_frame+20`` for values >= 236 (affects the low-pass and shot noise only).
    ``frame_index`` int32 ``[B,N]`` / ``value_map`` uint8 ``[B,256]``: the dataset's pause gather and HDR/LDR degrade fused
    into the pass, as in ``frames_to_voxel`` (``frames`` is then the raw stack ``[B,M,H,W]``)."""
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    if not frames.is_cuda or frames.dtype != torch.uint8:
        raise _lib.V2VError(-1, "frames must be a CUDA uint8 tensor (no CPU fallback)")
    frames = frames.contiguous()
    dev = frames.device
    B, N, H, W = frames.shape
    M = N
    fidx_t = vmap_t = None
    if frame_index is not None:
        fidx_t = torch.as_tensor(frame_index).to(device=dev, dtype=torch.int32)
        if fidx_t.dim() == 1:
            fidx_t = fidx_t.unsqueeze(0).expand(B, -1)
        if fidx_t.dim() != 2 or fidx_t.shape[0] != B:
            raise ValueError("frame_index must be [N] or [B,N]")
        fidx_t = fidx_t.contiguous()
        N = int(fidx_t.shape[1])
    if value_map is not None:
        vmap_t = torch.as_tensor(value_map).to(device=dev, dtype=torch.uint8)
        if vmap_t.dim() == 1:
            vmap_t = vmap_t.unsqueeze(0).expand(B, -1)
        if tuple(vmap_t.shape) != (B, 256):
            raise ValueError("value_map must be [256] or [B,256] uint8")
        vmap_t = vmap_t.contiguous()
    group = num_bins * frames_per_bin
    if (N - 1) % group != 0:
        raise AssertionError(f"(N-1)={N - 1} must be a multiple of num_bins*frames_per_bin={group}")
    T = (N - 1) // group
    per_interval = np.ndim(pos_thres) == 4          # [B,N-1,H,W]: maps re-drawn on every frame (spatial_temporal_independent)
    tshape = (B, N - 1, H, W) if per_interval else (B, H, W)
    pos_t = _as_dev(pos_thres, dev, torch.float64, tshape, "pos_thres")
    neg_t = _as_dev(neg_thres, dev, torch.float64, tshape, "neg_thres")
    if per_interval and noise == "philox":
        raise NotImplementedError("per-frame threshold maps come with host-drawn fields (noise 'none' or 'explicit')")
    nr_t = _as_dev(noise_rate, dev, torch.float32, (B, H, W), "noise_rate")
    lr_t = _as_dev(leak_randn, dev, torch.float64, (B, N - 1, H, W), "leak_randn")
    ps_t = _as_dev(pos_shot, dev, torch.int32, (B, N - 1, H, W), "pos_shot")
    ns_t = _as_dev(neg_shot, dev, torch.int32, (B, N - 1, H, W), "neg_shot")
    lut_t = torch.from_numpy(np.ascontiguousarray(v2e_log_lut() if lut is None else lut, dtype=np.float32)).to(dev)
    vox = torch.empty((B, T, num_bins, H, W), dtype=torch.float32, device=dev)
    stats_t = torch.zeros((B, 2), dtype=torch.int64, device=dev) if with_stats else None

    d = _lib.V2eDesc()
    d.B, d.N, d.H, d.W = B, N, H, W
    d.num_bins, d.frames_per_bin = num_bins, frames_per_bin
    d.noise_mode = _NOISE[noise]
    d.state_f32 = int(cutoff_hz <= 0 and leak_rate_hz <= 0)        # dtype rule, SURVEY Appendix A.2 step 5
    d.fps = float(fps)
    d.cutoff_hz, d.leak_rate_hz = float(cutoff_hz), float(leak_rate_hz)
    d.shot_noise_rate_hz, d.leak_jitter_fraction = float(shot_noise_rate_hz), float(leak_jitter_fraction)
    d.frames, d.lut = _ptr(frames), _ptr(lut_t)
    d.pos_thres, d.neg_thres, d.noise_rate = _ptr(pos_t), _ptr(neg_t), _ptr(nr_t)
    d.thres_per_interval = int(per_interval)
    d.leak_randn, d.pos_shot, d.neg_shot = _ptr(lr_t), _ptr(ps_t), _ptr(ns_t)
    d.pos_thres_nominal, d.neg_thres_nominal = float(pos_thres_nominal), float(neg_thres_nominal)
    d.seed, d.clip_index_base = int(seed) & 0xFFFFFFFFFFFFFFFF, int(clip_index_base)
    d.voxel, d.stats = _ptr(vox), _ptr(stats_t)
    d.frame_index, d.raw_frames_per_clip, d.value_map = _ptr(fidx_t), (M if fidx_t is not None else 0), _ptr(vmap_t)
    d.kernel_flags = int(kernel_flags) | _env_kernel_flags() | (_lib.V2E_FLAG_U8_INTENSITY if u8_intensity else 0)
    s = torch.cuda.current_stream(dev)
    lib = _lib.load()
    scales = None
    with torch.cuda.device(dev):
        if noise == "philox" and shot_noise_rate_hz > 0:
            scales = torch.empty((2, B, N - 1), dtype=torch.float64, device=dev)
            _lib.check(lib.v2v_v2e_shot_scales(C.byref(d), C.c_void_p(scales[0].data_ptr()),
                                               C.c_void_p(scales[1].data_ptr()), C.c_void_p(s.cuda_stream)))
            d.shot_pos_scale, d.shot_neg_scale = _ptr(scales[0]), _ptr(scales[1])
        fields = None
        if return_fields and noise == "philox":          # audit hook: what the generator drew for this very call
            fl = torch.zeros((B, N - 1, H, W), dtype=torch.float64, device=dev)
            fp = torch.zeros((B, N - 1, H, W), dtype=torch.int32, device=dev)
            fn = torch.zeros_like(fp)
            _lib.check(lib.v2v_v2e_philox_fields(C.byref(d), _ptr(fl), _ptr(fp), _ptr(fn), C.c_void_p(s.cuda_stream)))
            fields = {"leak_randn": fl, "pos_shot": fp, "neg_shot": fn}
        _lib.check(lib.v2v_v2e_frames_to_voxel(C.byref(d), C.c_void_p(s.cuda_stream)))
    return {"voxel": vox, "stats": stats_t, "fields": fields, "shot_scales": scales}


def video_to_voxel(video, FPS, threshold_model, thres_mean_mean, thres_mean_std, thres_diff_mean, thres_diff_std,
                   cutoff_hz, leak_rate_hz, refractory_period_s, shot_noise_rate_hz, leak_jitter_fraction,
                   noise_rate_cov_decades, seed, *, rng: str = "numpy", device="cuda", lut=None):
    """Same contract as the reference's ``video_to_voxel`` (data/v2v_core_v2e.py:556-581):
    ``video`` ``[N,H,W]`` with integer values 0..255 -> float64 ``[N-1,H,W]``.  Like the reference, a uint8 array takes
    uint8 arithmetic in ``rescale_intensity_frame`` (:190: ``new_frame+20`` wraps for values >= 236, which changes the
    low-pass constant and the shot-noise rate of bright pixels); any other dtype does not wrap.

    rng="numpy": every random field is drawn from the global legacy NumPy stream
    in the reference's order (seeded by ``seed`` exactly like the reference's
    constructor, :312-314) and replayed on the GPU — bit-identical to the
    reference.  rng="philox": threshold / noise-rate maps from NumPy, per-frame
    leak jitter and shot noise generated in the kernel.
    """
    if refractory_period_s and refractory_period_s > 0:
        raise TypeError("refractory_period_s > 0: the reference's branch calls np.clip(x, a_max=...) "
                        "and raises TypeError (SURVEY §4); not implemented")
    vid = np.asarray(video)
    u8 = vid.dtype == np.uint8
    if not u8:
        if not np.array_equal(vid, np.clip(np.rint(vid), 0, 255)):
            raise ValueError("video must hold integer values in 0..255")
        vid8 = vid.astype(np.uint8)
    else:
        vid8 = vid
    N, H, W = vid8.shape
    if seed is not None:                                                        # :312-314
        np.random.seed(seed)
    per_frame = threshold_model == PER_FRAME_MODEL
    if per_frame:
        if rng != "numpy":
            raise NotImplementedError("threshold_model 'spatial_temporal_independent' redraws two full-frame maps from the "
                                      "NumPy stream before every frame: rng='numpy' only")
        threshold_maps(threshold_model, thres_mean_mean, thres_mean_std, 0, 0, (H, W))     # :417-421 on frame 0 (then _init draws again)
    pos, neg = threshold_maps(threshold_model, thres_mean_mean, thres_mean_std, thres_diff_mean, thres_diff_std,
                              (H, W))
    nrate = noise_rate_map(noise_rate_cov_decades, (H, W))
    pos_nom = thres_mean_mean + thres_diff_mean / 2                             # :297-298
    neg_nom = thres_mean_mean - thres_diff_mean / 2
    kw = dict(noise="none")
    if rng == "numpy":
        leak_r = pos_s = neg_s = None
        pos_f = np.empty((N - 1, H, W)) if per_frame else None
        neg_f = np.empty((N - 1, H, W)) if per_frame else None
        if leak_rate_hz > 0 or shot_noise_rate_hz > 0 or per_frame:
            leak_r = np.zeros((N - 1, H, W)) if leak_rate_hz > 0 else None
            if shot_noise_rate_hz > 0:
                pos_s = np.zeros((N - 1, H, W), dtype=np.int32)
                neg_s = np.zeros((N - 1, H, W), dtype=np.int32)
                pos_pp, neg_pp = np.divide(pos_nom, pos), np.divide(neg_nom, neg)
            t_prev = 0.0
            for k in range(1, N):                       # per-frame draw order: (threshold maps,) leak randn, poisson x2
                t_k = k / FPS
                dt = t_k - t_prev
                t_prev = t_k
                if per_frame:                                                   # :417-421
                    pos_f[k - 1], neg_f[k - 1] = threshold_maps(threshold_model, thres_mean_mean, thres_mean_std, 0, 0, (H, W))
                    if shot_noise_rate_hz > 0:
                        pos_pp, neg_pp = np.divide(pos_nom, pos_f[k - 1]), np.divide(neg_nom, neg_f[k - 1])
                if leak_rate_hz > 0:
                    leak_r[k - 1] = np.random.randn(H, W)                       # :201
                if shot_noise_rate_hz > 0:                                      # :90-103
                    inten01 = ((vid8[k] + np.uint8(20)) if u8 else (vid8[k].astype(np.float64) + 20)) / 275.
                    fac = 1 - (1 - 0.25) * inten01
                    pf = fac * pos_pp
                    pf = pf / np.mean(pf)
                    nf = fac * neg_pp
                    nf = nf / np.mean(nf)
                    sf = (shot_noise_rate_hz / 2) * dt
                    pos_s[k - 1] = np.random.poisson(pf * sf)
                    neg_s[k - 1] = np.random.poisson(nf * sf)
        kw = dict(noise="explicit",
                  leak_randn=None if leak_r is None else leak_r[None],
                  pos_shot=None if pos_s is None else pos_s[None],
                  neg_shot=None if neg_s is None else neg_s[None])
    elif rng == "philox":
        # (the reference seeds NumPy only when a seed is given, :312-314: without one every call draws fresh noise)
        kw = dict(noise="philox", seed=int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed)
    else:
        raise ValueError("rng must be 'numpy' or 'philox'")
    if N < 2:
        return np.zeros((0, H, W))
    frames = torch.from_numpy(np.ascontiguousarray(vid8)).to(device)
    if per_frame:
        pos, neg = pos_f, neg_f
    out = frames_to_voxel_v2e(frames, pos[None], neg[None], fps=FPS, cutoff_hz=cutoff_hz, leak_rate_hz=leak_rate_hz,
                              shot_noise_rate_hz=shot_noise_rate_hz, leak_jitter_fraction=leak_jitter_fraction,
                              noise_rate=nrate[None], pos_thres_nominal=pos_nom, neg_thres_nominal=neg_nom,
                              lut=lut, u8_intensity=u8, **kw)
    return out["voxel"][0, :, 0].to(torch.float64).cpu().numpy()
