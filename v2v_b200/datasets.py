"""Dataset-level boundary of the frame path: ``imgs_to_voxels`` and batch packing.

Mirrors ``WebvidDatasetV2.imgs_to_voxels`` (reference data/v2v_datasets.py:363-410)
and the tensor packaging of ``__getitem__`` (:328-356).  Video decoding, cropping,
pausing and degradation stay in the reference (CPU video I/O, out of scope).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .esim import EventEmulator, default_rng_mode, frames_to_voxel


def sample_v2e_params(configs, pos_thres=None, neg_thres=None, rs=np.random) -> Dict[str, float]:
    """The simulator-parameter sampling law of data/v2v_datasets.py:368-386.

    Consumes ``rs`` (default: the global legacy NumPy stream) in the reference's
    order.  ``configs`` is the object holding ``threshold_range``,
    ``max_thres_pos_neg_gap``, ``base_noise_std_range``,
    ``hot_pixel_fraction_range``, ``hot_pixel_std_range``,
    ``use_fixed_thresholds``, ``scale_noise_strength``, ``put_noise_external``.
    """
    if not configs.use_fixed_thresholds:
        thres_1 = rs.uniform(*configs.threshold_range)
        gap = rs.uniform(1, configs.max_thres_pos_neg_gap)
        thres_2 = thres_1 * gap
        if rs.rand() > 0.5:
            pos_thres, neg_thres = thres_1, thres_2
        else:
            pos_thres, neg_thres = thres_2, thres_1
    base_noise_std = rs.uniform(*configs.base_noise_std_range)
    hot_pixel_fraction = rs.uniform(*configs.hot_pixel_fraction_range)
    hot_pixel_std = rs.uniform(*configs.hot_pixel_std_range)
    if configs.scale_noise_strength and not configs.put_noise_external:
        base_noise_std = base_noise_std * pos_thres
        hot_pixel_std = hot_pixel_std * pos_thres
    return {"pos_thres": pos_thres, "neg_thres": neg_thres, "base_noise_std": base_noise_std,
            "hot_pixel_fraction": hot_pixel_fraction, "hot_pixel_std": hot_pixel_std}


def sample_pause_indices(count: int, proba_pause_when_running: float, proba_pause_when_paused: float, rs=np.random):
    """The dataset's pause sequence (data/v2v_datasets.py:285-301): ``count`` frame indices into the raw clip,
    non-decreasing with repeats while "paused"; one ``rs.rand()`` per frame in the reference's order.
    Returns (img_idxes int32 [count], true_img_cnt) — ``true_img_cnt`` raw frames have to be decoded."""
    img_idxes, idx, is_pause = [], 0, False
    for _ in range(count):
        img_idxes.append(idx)
        if is_pause and rs.rand() > proba_pause_when_paused:
            is_pause = False
        elif not is_pause and rs.rand() < proba_pause_when_running:
            is_pause = True
        if not is_pause:
            idx += 1
    return np.asarray(img_idxes, dtype=np.int32), idx + 1


def degrade_value_map(kind: str, scale: float) -> np.ndarray:
    """uint8[256] map of the HDR / LDR degrade (data/v2v_datasets.py:473-483): the reference's own NumPy expression
    evaluated on every pixel value, so applying the map equals degrading the frames bit for bit.
    kind: "hdr" (scale ~ U(1,3)) or "ldr" (scale ~ U(0.3,1)); the caller draws ``scale`` like the reference does."""
    if kind not in ("hdr", "ldr"):
        raise NotImplementedError("Video degrade type not supported.")
    v = np.arange(256, dtype=np.uint8)
    return np.clip((v - 127.5) * scale + 127.5, 0, 255).astype(np.uint8)


def bgr_to_gray(img_stack: torch.Tensor) -> torch.Tensor:
    """CUDA uint8 ``[..., C>=3]`` -> uint8 ``[...]``: the dataset's ``bgr_to_gray`` (data/v2v_datasets.py:19-22)."""
    import ctypes as C
    from . import _lib
    if not img_stack.is_cuda or img_stack.dtype != torch.uint8 or img_stack.shape[-1] < 3:
        raise _lib.V2VError(-1, "bgr_to_gray needs a CUDA uint8 tensor [..., C>=3] (no CPU fallback)")
    img = img_stack.contiguous()
    gray = torch.empty(img.shape[:-1], dtype=torch.uint8, device=img.device)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().v2v_bgr_to_gray(C.c_void_p(img.data_ptr()), int(img.shape[-1]), C.c_void_p(gray.data_ptr()),
                                               int(gray.numel()), C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)))
    return gray


class ImgsToVoxelsMixin:
    """GPU override of ``WebvidDatasetV2.imgs_to_voxels``.

    ``class WebvidDatasetV2B200(ImgsToVoxelsMixin, WebvidDatasetV2): pass`` and
    point the yaml ``class_name`` at it (the reference resolves datasets by name,
    data/data_interface.py:6-20).  Everything else — decode, crop, pause,
    degrade, packing — is inherited unchanged.  With CUDA in DataLoader workers
    use ``num_workers: 0`` or the spawn start method.
    """

    v2v_rng: Optional[str] = None      # None -> env V2V_B200_RNG or "philox"
    v2v_device = "cuda"

    def imgs_to_voxels(self, imgs, num_bins, frames_per_bin, FPS, pos_thres=None, neg_thres=None):
        N, H, W = imgs.shape
        assert (N - 1) % (num_bins * frames_per_bin) == 0                      # :365
        frame_cnt = (N - 1) // (num_bins * frames_per_bin)
        p = sample_v2e_params(self, pos_thres, neg_thres)
        all_voxels = EventEmulator(
            pos_thres=p["pos_thres"], neg_thres=p["neg_thres"], base_noise_std=p["base_noise_std"],
            hot_pixel_fraction=p["hot_pixel_fraction"], hot_pixel_std=p["hot_pixel_std"],
            put_noise_external=self.put_noise_external, seed=None,
            rng=self.v2v_rng or default_rng_mode(), device=self.v2v_device,
        ).video_to_voxel(np.ascontiguousarray(imgs))
        all_voxels = all_voxels.reshape((frame_cnt, num_bins, frames_per_bin, H, W))      # :399
        return p, all_voxels.sum(axis=2)                                                  # :400


class V2VVoxelizer(ImgsToVoxelsMixin):
    """Stand-alone holder of the path's configuration (the subset of
    ``WebvidDatasetV2.load_configs`` that the hot path reads, data/v2v_datasets.py:40-60)
    with the batched GPU-resident entry the train loop can consume directly."""

    def __init__(self, configs: Optional[dict] = None, device="cuda", rng: Optional[str] = None):
        configs = configs or {}
        self.num_bins = configs.get("num_bins", 5)
        self.frames_per_bin = configs.get("frames_per_bin", 1)
        self.frames_per_img = self.num_bins * self.frames_per_bin
        self.threshold_range = configs.get("threshold_range", [0.05, 2])
        self.max_thres_pos_neg_gap = configs.get("max_thres_pos_neg_gap", 1.5)
        self.base_noise_std_range = configs.get("base_noise_std_range", [0, 0.2])
        self.hot_pixel_fraction_range = configs.get("hot_pixel_fraction_range", [0, 0.001])
        self.hot_pixel_std_range = configs.get("hot_pixel_std_range", [0, 0.2])
        self.put_noise_external = configs.get("put_noise_external", False)
        self.scale_noise_strength = configs.get("scale_noise_strength", False)
        self.use_fixed_thresholds = configs.get("use_fixed_thresholds", False)
        self.output_additional_frame = configs.get("output_additional_frame", False)
        self.v2v_device = device
        self.v2v_rng = rng

    def sample_batch_params(self, batch: int, fixed: Optional[Sequence] = None, rs=np.random):
        """One parameter draw per clip, in clip order (as ``batch`` successive ``__getitem__`` calls would)."""
        ps = []
        for b in range(batch):
            f = fixed[b] if fixed is not None else (None, None)
            ps.append(sample_v2e_params(self, f[0], f[1], rs))
        return ps

    def batch_to_tensors(self, frames: torch.Tensor, params: Optional[Sequence[dict]] = None, *, seed: int = 0,
                         clip_index_base: int = 0, pad_multiple: int = 0, with_stats: bool = False, out=None,
                         frame_index=None, value_map=None):
        """CUDA uint8 ``[B,N,H,W]`` gray clips -> the train batch dict on the GPU.  With ``frame_index`` ``[B,N]`` the input
        is the raw decoded stack ``[B,M,H,W]`` and the pause gather happens inside the kernel; ``value_map`` ``[B,256]``
        applies the HDR/LDR degrade (``sample_pause_indices``, ``degrade_value_map``).

        Returns {"events": float32 [B,T,bins,H,W], "frame": float32 [B,T(+1),1,H,W] in [0,1],
        "v2e_params": list of dicts[, "stats": int64 [B,2]]} — the layout
        ``default_collate`` produces from data/v2v_datasets.py:351-356 and
        train.py:79-83 / model/train_utils.py:318 consume.  Noise comes from the
        in-kernel Philox generator keyed by (seed, clip_index_base + b).
        """
        B = frames.shape[0]
        if params is None:
            params = self.sample_batch_params(B)
        col = lambda k: np.array([p[k] for p in params], dtype=np.float64)
        o = frames_to_voxel(
            frames, col("pos_thres"), col("neg_thres"), num_bins=self.num_bins, frames_per_bin=self.frames_per_bin,
            noise="philox", base_noise_std=col("base_noise_std"), hot_pixel_fraction=col("hot_pixel_fraction"),
            hot_pixel_std=col("hot_pixel_std"), put_noise_external=self.put_noise_external, seed=seed,
            clip_index_base=clip_index_base, pad_multiple=pad_multiple, with_stats=with_stats, out=out,
            frame_out="frames+first" if self.output_additional_frame else "frames", frame_index=frame_index, value_map=value_map)
        batch = {"events": o.voxel, "frame": o.frames, "v2e_params": list(params)}
        if with_stats:
            batch["stats"] = o.stats
        if o.padded is not None:
            batch["events_padded"] = o.padded
        return batch
