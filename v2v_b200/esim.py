"""ESIM-style video -> voxel on B200: host side of the frame path.

Two layers:

* ``frames_to_voxel`` — the batched, GPU-resident API (CUDA uint8 ``[B,N,H,W]``
  in, CUDA float32 ``[B,T,bins,H,W]`` out) that the benchmark, the dataset shim
  and multi-GPU sharding use.  One fused kernel launch per call.
* ``EventEmulator`` — the reference's class, same constructor and
  ``video_to_voxel(video) -> float64 [N-1,H,W]`` (reference
  data/v2v_core_esim.py:6-69; call sites data/v2v_datasets.py:388-396,
  scripts/visualize_esim_sample.py:178-186).

Both call the C ABI (include/v2v_b200.h) through ctypes; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib

_NOISE = {"none": _lib.NOISE_NONE, "explicit": _lib.NOISE_EXPLICIT, "philox": _lib.NOISE_PHILOX}


def esim_log_lut() -> np.ndarray:
    """float64[256]: log(0.001 + reverse_gamma(v)/255) for v = 0..255.

    Evaluated on the host with the reference's NumPy expressions
    (data/v2v_core_esim.py:4 and :34) so that ``lut[video]`` equals the
    reference's ``log_imgs`` bit for bit; the kernel only ever indexes it.
    """
    v = np.arange(256, dtype=np.uint8)
    return np.log(0.001 + ((v / 255) ** 2.2 * 255) / 255.0)


_lut_cache = {}


def _device_lut(device: torch.device, lut: Optional[np.ndarray]) -> torch.Tensor:
    if lut is not None:
        arr = np.ascontiguousarray(lut, dtype=np.float64)
        if arr.shape != (256,):
            raise ValueError("lut must have 256 float64 entries")
        return torch.from_numpy(arr).to(device)
    key = (device.type, device.index)
    if key not in _lut_cache:
        _lut_cache[key] = torch.from_numpy(esim_log_lut()).to(device)
    return _lut_cache[key]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _as_dev(x, device, dtype, shape=None, name="tensor"):
    """Return a contiguous device tensor of ``dtype`` (None passes through)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    x = x.to(device=device, dtype=dtype).contiguous()
    if shape is not None and tuple(x.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(x.shape)}, expected {tuple(shape)}")
    return x


@dataclass
class EsimOutput:
    voxel: torch.Tensor                       # [B,T,bins,H,W] float32 (a view when padded)
    frames: Optional[torch.Tensor] = None     # [B,T(+1),1,H,W] float32 in [0,1]
    stats: Optional[torch.Tensor] = None      # [B,2] int64: positive / negative event totals
    potential: Optional[torch.Tensor] = None  # [B,H,W] float64 membrane potential after the last frame
    padded: Optional[torch.Tensor] = None     # [B,T,bins,Hp,Wp] storage when pad_multiple is used


def frames_to_voxel(frames: torch.Tensor, pos_thres, neg_thres, *, num_bins: int = 5, frames_per_bin: int = 1,
                    noise: str = "none", base_noise_std=0.0, hot_pixel_fraction=0.0, hot_pixel_std=0.0,
                    put_noise_external: bool = False, u0=None, hot_noise=None, base_gauss=None,
                    seed: int = 0, clip_index_base: int = 0, potential_in=None, return_potential: bool = False,
                    frame_out: Optional[str] = None, with_stats: bool = False, pad_multiple: int = 0,
                    out: Optional[torch.Tensor] = None, lut: Optional[np.ndarray] = None,
                    stream: Optional[torch.cuda.Stream] = None, frame_index=None, value_map=None,
                    kernel_flags: int = 0) -> EsimOutput:
    """Simulate ``B`` clips in one launch.

    frames: CUDA uint8 ``[B,N,H,W]`` (or ``[N,H,W]``).  pos_thres / neg_thres:
    scalar, ``[B]`` (per clip, the ESIM core) or ``[B,H,W]`` (per-pixel maps).
    noise: "none" | "explicit" (u0, hot_noise ``[B,H,W]``, base_gauss
    ``[B,N-1,H,W]`` float64: the reference's random fields) | "philox"
    (in-kernel generator keyed by ``seed``: Philox4x32-10 root + one 64-bit LCG stream per pixel group; clip ``b`` uses stream
    ``clip_index_base + b``).  ``num_bins*frames_per_bin`` must divide ``N-1``
    (data/v2v_datasets.py:365).  frame_out: None | "frames" (frames
    ``(t+1)*bins*fpb``) | "frames+first" (``t*bins*fpb``, t<=T;
    output_additional_frame).  pad_multiple: allocate the voxel with H and W
    rounded up (the consumer's /16 padding, model/train_utils.py:322-326); pads
    are zero and the returned ``voxel`` is the unpadded view.

    Frame-side packing fused into the pass (data/v2v_datasets.py:285-311,473-483): ``frame_index`` int32 ``[B,N]`` makes
    frame ``n`` of clip ``b`` the raw frame ``frames[b, frame_index[b,n]]`` (the dataset's pause gather; ``frames`` is
    then the raw stack ``[B,M,H,W]``); ``value_map`` uint8 ``[B,256]`` is applied to every pixel before the simulation and
    the ``frame_out`` (``degrade_value_map`` builds the HDR/LDR degrade).

    ``kernel_flags``: ``_lib.ESIM_FLAG_*`` (explicit kernel selection for tests and tuning sweeps; results never depend
    on it).  ``stream``: the launch and every allocation / conversion it depends on are ordered on that stream.
    """
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
    if N < 1 or (N - 1) % group != 0:
        raise AssertionError(f"(N-1)={N - 1} must be a multiple of num_bins*frames_per_bin={group}")
    T = (N - 1) // group
    if noise not in _NOISE:
        raise ValueError(f"noise must be one of {sorted(_NOISE)}")

    def thr(x, name):
        if (isinstance(x, torch.Tensor) and x.device == dev and x.dtype == torch.float64 and x.dim() == 1
                and x.shape[0] == B and x.is_contiguous()):
            return x, _lib.THRES_PER_CLIP                      # already in place: no host work
        t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
        t = t.to(device=dev, dtype=torch.float64)
        if t.dim() == 0:
            t = t.expand(B)
        if t.dim() == 1:
            if t.shape[0] != B:
                raise ValueError(f"{name} has {t.shape[0]} entries for {B} clips")
            return t.contiguous(), _lib.THRES_PER_CLIP
        if t.dim() == 2:
            t = t.unsqueeze(0).expand(B, H, W)
        if tuple(t.shape) != (B, H, W):
            raise ValueError(f"{name} must be scalar, [B] or [B,H,W]")
        return t.contiguous(), _lib.THRES_PER_PIXEL

    pos_t, pm = thr(pos_thres, "pos_thres")
    neg_t, nm = thr(neg_thres, "neg_thres")
    if pm != nm:
        if pm == _lib.THRES_PER_CLIP:
            pos_t = pos_t.view(B, 1, 1).expand(B, H, W).contiguous()
        else:
            neg_t = neg_t.view(B, 1, 1).expand(B, H, W).contiguous()
        pm = _lib.THRES_PER_PIXEL

    def per_clip(x):
        if (isinstance(x, torch.Tensor) and x.device == dev and x.dtype == torch.float64 and x.dim() == 1
                and x.shape[0] == B and x.is_contiguous()):
            return x
        t = torch.as_tensor(np.asarray(x, dtype=np.float64)) if not isinstance(x, torch.Tensor) else x
        t = t.to(device=dev, dtype=torch.float64)
        return (t.expand(B) if t.dim() == 0 else t).contiguous()

    std_t = per_clip(base_noise_std) if noise != "none" else None
    frac_t = per_clip(hot_pixel_fraction) if noise == "philox" else None
    hstd_t = per_clip(hot_pixel_std) if noise == "philox" else None
    u0_t = _as_dev(u0, dev, torch.float64, (B, H, W), "u0")
    hot_t = _as_dev(hot_noise, dev, torch.float64, (B, H, W), "hot_noise")
    g_t = _as_dev(base_gauss, dev, torch.float64, (B, N - 1, H, W), "base_gauss")
    pin_t = _as_dev(potential_in, dev, torch.float64, (B, H, W), "potential_in")
    pout_t = torch.empty((B, H, W), dtype=torch.float64, device=dev) if return_potential else None

    Hp, Wp = H, W
    if pad_multiple and pad_multiple > 1:
        Hp = -(-H // pad_multiple) * pad_multiple
        Wp = -(-W // pad_multiple) * pad_multiple
    if out is None:
        alloc = torch.zeros if (Hp, Wp) != (H, W) else torch.empty
        store = alloc((B, T, num_bins, Hp, Wp), dtype=torch.float32, device=dev)
    else:
        store = out
        if (not store.is_cuda or store.dtype != torch.float32 or not store.is_contiguous()
                or tuple(store.shape) != (B, T, num_bins, Hp, Wp)):
            raise ValueError(f"out must be a contiguous CUDA float32 tensor of shape {(B, T, num_bins, Hp, Wp)}")
    fmode = {None: 0, "frames": 1, "frames+first": 2}[frame_out]
    fr_t = None
    if fmode:
        fr_t = torch.empty((B, T + (1 if fmode == 2 else 0), 1, H, W), dtype=torch.float32, device=dev)
    stats_t = torch.zeros((B, 2), dtype=torch.int64, device=dev) if with_stats else None
    lut_t = _device_lut(dev, lut)

    d = _lib.EsimDesc()
    d.B, d.N, d.H, d.W = B, N, H, W
    d.num_bins, d.frames_per_bin = num_bins, frames_per_bin
    d.noise_mode = _NOISE[noise]
    d.put_noise_external = int(bool(put_noise_external))
    d.threshold_mode = pm
    d.frame_out_mode = fmode
    d.frames, d.lut = _ptr(frames), _ptr(lut_t)
    d.pos_thres, d.neg_thres, d.base_noise_std = _ptr(pos_t), _ptr(neg_t), _ptr(std_t)
    d.u0, d.hot_noise, d.base_gauss = _ptr(u0_t), _ptr(hot_t), _ptr(g_t)
    d.hot_pixel_fraction, d.hot_pixel_std = _ptr(frac_t), _ptr(hstd_t)
    d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    d.clip_index_base = int(clip_index_base)
    d.potential_in, d.potential_out = _ptr(pin_t), _ptr(pout_t)
    d.voxel = _ptr(store)
    d.voxel_row_stride, d.voxel_plane_stride = Wp, Hp * Wp
    d.frame_out, d.stats = _ptr(fr_t), _ptr(stats_t)
    d.frame_index, d.raw_frames_per_clip, d.value_map = _ptr(fidx_t), (M if fidx_t is not None else 0), _ptr(vmap_t)
    d.kernel_flags = int(kernel_flags) | _env_kernel_flags()

    s = stream if stream is not None else torch.cuda.current_stream(dev)
    if stream is not None:       # conversions, zero fills and H2D copies above were enqueued on the current stream
        stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_esim_frames_to_voxel(C.byref(d), C.c_void_p(s.cuda_stream)))
    if stream is not None:       # tensors created on the current stream but consumed on `stream`
        for t in (frames, pos_t, neg_t, std_t, frac_t, hstd_t, u0_t, hot_t, g_t, pin_t, lut_t, fidx_t, vmap_t):
            if t is not None:
                t.record_stream(s)
    vox = store[..., :H, :W] if (Hp, Wp) != (H, W) else store
    return EsimOutput(voxel=vox, frames=fr_t, stats=stats_t, potential=pout_t,
                      padded=store if (Hp, Wp) != (H, W) else None)


def _env_kernel_flags() -> int:
    """Tuning knobs of the sweep tools (read here, in the Python host, per call; the C library reads no environment)."""
    f = 0
    g = os.environ.get("V2V_ESIM_GEOM")
    if g:
        f |= _lib.esim_flag_geom(int(g))
    if os.environ.get("V2V_ESIM_GENERIC") == "1":
        f |= _lib.ESIM_FLAG_GENERIC
    if os.environ.get("V2V_ESIM_STAGED") == "1":
        f |= _lib.ESIM_FLAG_STAGED
    return f


def philox_fields(n_frames: int, height: int, width: int, *, base_noise_std, hot_pixel_fraction, hot_pixel_std,
                  seed: int = 0, clip_index_base: int = 0, device="cuda"):
    """The random fields a ``noise="philox"`` run draws, as float64 tensors (u0 [B,H,W], hot_noise [B,H,W],
    base_noise [B,N-1,H,W] already scaled by base_noise_std).  Replaying them through ``noise="explicit"`` with
    ``base_noise_std=1`` reproduces the Philox run bit for bit — the hook the tests use to check the in-kernel
    generator path against the CPU oracle."""
    dev = torch.device(device)
    to = lambda x: torch.as_tensor(np.atleast_1d(np.asarray(x, dtype=np.float64))).to(dev).contiguous()
    std_t, frac_t, hstd_t = to(base_noise_std), to(hot_pixel_fraction), to(hot_pixel_std)
    B = std_t.shape[0]
    u0 = torch.empty((B, height, width), dtype=torch.float64, device=dev)
    hot = torch.empty_like(u0)
    bn = torch.empty((B, n_frames - 1, height, width), dtype=torch.float64, device=dev)
    d = _lib.EsimDesc()
    d.B, d.N, d.H, d.W = B, n_frames, height, width
    d.num_bins = d.frames_per_bin = 1
    d.noise_mode = _lib.NOISE_PHILOX
    d.base_noise_std, d.hot_pixel_fraction, d.hot_pixel_std = _ptr(std_t), _ptr(frac_t), _ptr(hstd_t)
    d.seed, d.clip_index_base = int(seed) & 0xFFFFFFFFFFFFFFFF, int(clip_index_base)
    s = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_esim_philox_fields(C.byref(d), _ptr(u0), _ptr(hot), _ptr(bn), C.c_void_p(s.cuda_stream)))
    return u0, hot, bn


def draw_reference_randomness(n_frames, height, width, hot_pixel_fraction, hot_pixel_std, rs=np.random):
    """Consume ``rs`` (default: the global legacy NumPy stream, as the reference
    does) in the reference's order and return (u0, hot_noise, base_gauss).

    data/v2v_core_esim.py:29 (rand), :37 (rand), :38 (randn), :44 (randn per
    interval).  Feeding these fields to the kernel in "explicit" mode reproduces
    the reference bit for bit under the same ``np.random.seed``.
    """
    u0 = rs.rand(height, width)
    mask = rs.rand(height, width) < hot_pixel_fraction
    hot = np.where(mask, hot_pixel_std * rs.randn(height, width), 0)
    g = rs.randn(n_frames - 1, height, width)
    return u0, hot, g


def default_rng_mode() -> str:
    return os.environ.get("V2V_B200_RNG", "philox")


class EventEmulator(object):
    """Drop-in for the reference's ``EventEmulator`` (data/v2v_core_esim.py:6-69).

    Same constructor arguments and ``video_to_voxel`` contract.  Two extra
    keyword-only knobs:

    rng: "numpy"  — draw the random fields from the global legacy ``np.random``
                    stream in the reference's order on the host and replay them
                    on the GPU: bit-identical to the reference for the same
                    ``np.random.seed`` (parity mode; the host RNG is the
                    bottleneck).
         "philox" — generate them in the kernel (throughput mode, default;
                    override with env V2V_B200_RNG).  The Philox key is
                    ``seed`` if given, else one draw from ``np.random`` so that
                    ``np.random.seed`` still makes a run reproducible.
    device: CUDA device the simulation runs on.
    """

    def __init__(self, pos_thres: float = 0.2, neg_thres: float = 0.2, base_noise_std: float = 0.1,
                 hot_pixel_fraction: float = 0.001, hot_pixel_std: float = 0.1, put_noise_external: bool = False,
                 seed: int = None, *, rng: Optional[str] = None, device="cuda"):
        self.pos_threshold = pos_thres
        self.neg_threshold = neg_thres
        self.base_noise_std = base_noise_std
        self.hot_pixel_fraction = hot_pixel_fraction
        self.hot_pixel_std = hot_pixel_std
        self.put_noise_external = put_noise_external
        self.seed = seed
        if not (pos_thres > 0 and neg_thres > 0):
            raise ValueError("pos_thres and neg_thres must be positive")
        self.rng = rng or default_rng_mode()
        if self.rng not in ("numpy", "philox"):
            raise ValueError("rng must be 'numpy' or 'philox'")
        self.device = torch.device(device)
        self.potential = None
        self.last_stats = None

    def video_to_voxel(self, video, lut: Optional[np.ndarray] = None):
        """uint8 ``[N,H,W]`` -> float64 ``[N-1,H,W]`` signed event counts per interval."""
        video = np.asarray(video)
        if video.ndim != 3:
            raise ValueError("video must be [N,H,W]")
        if video.dtype != np.uint8:
            raise TypeError("video must be uint8 (the reference's frames always are, data/v2v_datasets.py:19-22)")
        N, H, W = video.shape
        kw = {}
        if self.rng == "numpy":
            u0, hot, g = draw_reference_randomness(N, H, W, self.hot_pixel_fraction, self.hot_pixel_std)
            kw = dict(noise="explicit", u0=u0[None], hot_noise=hot[None], base_gauss=g[None])
        else:
            seed = self.seed if self.seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
            kw = dict(noise="philox", seed=seed, hot_pixel_fraction=self.hot_pixel_fraction,
                      hot_pixel_std=self.hot_pixel_std)
        if N < 2:
            return np.zeros((0, H, W))
        frames = torch.from_numpy(np.ascontiguousarray(video)).to(self.device)
        out = frames_to_voxel(frames, self.pos_threshold, self.neg_threshold, num_bins=1, frames_per_bin=1,
                              base_noise_std=self.base_noise_std, put_noise_external=self.put_noise_external,
                              return_potential=True, with_stats=True, lut=lut, **kw)
        self.potential = out.potential[0].cpu().numpy()
        self.last_stats = out.stats[0].cpu().numpy()
        return out.voxel[0, :, 0].to(torch.float64).cpu().numpy()
