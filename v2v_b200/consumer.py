"""Consumer side of the voxel tensor: ``normalize_batch_voxel`` and the per-bin sums of the
statistics vector, on CUDA float32 voxels.

``normalize_batch_voxel`` mirrors model/train_utils.py:147-166 (same name, argument and result).  The reference finds the
1 % / 99 % order statistics of every clip with ``torch.kthvalue`` over ``T*C*H*W`` elements; voxels of the frame path are
small integers, so one histogram pass (``v2v_voxel_value_hist``) gives the same k-th values exactly, and when both come
out <= 1 — every clip whose event rate is below 1 % per polarity, and every clip whose 99th percentile is a single
event — the normalisation is the identity and no second pass runs.  Non-integer voxels (external noise, cached voxels
with Gaussian noise) take the reference's ``torch.kthvalue`` for the order statistics and the same scale kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_K = 255


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def voxel_value_hist(voxel: torch.Tensor) -> torch.Tensor:
    """int64 ``[B, 512]``: counts of the integer values -255..255 per clip (index v+255) and, at index 511, of every other
    value."""
    if not voxel.is_cuda or voxel.dtype != torch.float32 or not voxel.is_contiguous():
        raise _lib.V2VError(-1, "voxel must be a contiguous CUDA float32 tensor (no CPU fallback)")
    B = voxel.shape[0]
    hist = torch.zeros((B, 2 * _K + 2), dtype=torch.int64, device=voxel.device)
    if voxel.numel():
        with torch.cuda.device(voxel.device):
            _lib.check(_lib.load().v2v_voxel_value_hist(_p(voxel), B, voxel.numel() // B, _p(hist),
                                                        C.c_void_p(torch.cuda.current_stream(voxel.device).cuda_stream)))
    return hist


def kth_from_hist(hist_row: np.ndarray, k: int) -> float:
    """k-th smallest value (1-indexed, like torch.kthvalue) of a clip whose exact histogram over -255..255 is given."""
    cum = np.cumsum(hist_row[: 2 * _K + 1])
    return float(int(np.searchsorted(cum, k, side="left")) - _K)


def normalize_batch_voxel(voxel: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    """model/train_utils.py:147-166 on a CUDA float32 ``[B,T,C,H,W]`` batch."""
    assert len(voxel.shape) == 5
    if not voxel.is_cuda or voxel.dtype != torch.float32:
        raise _lib.V2VError(-1, "voxel must be a CUDA float32 tensor (no CPU fallback)")
    B = voxel.shape[0]
    v = voxel if voxel.is_contiguous() else voxel.contiguous()
    n = v.numel() // max(B, 1)
    max_k, min_k = int(0.99 * n), int(0.01 * n)                     # :153-154
    if min_k < 1:
        raise RuntimeError("kthvalue(): selected number k out of range for dimension 1")      # what torch.kthvalue raises
    hist = voxel_value_hist(v).cpu().numpy()
    pos_max, neg_max = np.empty(B, dtype=np.float32), np.empty(B, dtype=np.float32)
    flat = None
    for b in range(B):
        if hist[b, 2 * _K + 1] == 0:
            pos_max[b], neg_max[b] = kth_from_hist(hist[b], max_k), -kth_from_hist(hist[b], min_k)
        else:                                                        # not integer valued: the reference's own way
            flat = v.reshape(B, -1) if flat is None else flat
            pos_max[b] = float(torch.kthvalue(flat[b], max_k).values)
            neg_max[b] = -float(torch.kthvalue(flat[b], min_k).values)
    pos_max, neg_max = np.maximum(pos_max, 1), np.maximum(neg_max, 1)       # :161-162
    if (pos_max == 1).all() and (neg_max == 1).all():
        return v if inplace else v.clone()                           # x / 1 == x: nothing to do
    out = v if inplace else v.clone()
    pm, nm = torch.from_numpy(pos_max).to(v.device), torch.from_numpy(neg_max).to(v.device)
    with torch.cuda.device(v.device):
        _lib.check(_lib.load().v2v_voxel_normalize(_p(out), B, n, _p(pm), _p(nm),
                                                   C.c_void_p(torch.cuda.current_stream(v.device).cuda_stream)))
    return out


def bin_abs_sums(voxel: torch.Tensor) -> torch.Tensor:
    """float64 ``[bins]``: sum of |voxel| per temporal bin of a CUDA float32 ``[..., bins, H, W]`` batch (exact for
    integer-valued voxels) — the per-bin event totals of the statistics vector."""
    if not voxel.is_cuda or voxel.dtype != torch.float32 or not voxel.is_contiguous() or voxel.dim() < 3:
        raise _lib.V2VError(-1, "voxel must be a contiguous CUDA float32 [..., bins, H, W] tensor")
    bins, plane = voxel.shape[-3], voxel.shape[-2] * voxel.shape[-1]
    groups = voxel.numel() // max(bins * plane, 1)
    sums = torch.zeros(bins, dtype=torch.float64, device=voxel.device)
    if voxel.numel():
        with torch.cuda.device(voxel.device):
            _lib.check(_lib.load().v2v_voxel_bin_abs_sums(_p(voxel), groups, bins, plane, _p(sums),
                                                          C.c_void_p(torch.cuda.current_stream(voxel.device).cuda_stream)))
    return sums


# ---- NER-Net's learned event representation (model/nernet/representation_modules.py:143-168, 228-248) ----------------
class _PutAccumulateBins(torch.autograd.Function):
    """vox.put_(idx + bin_stride*b, values[b], accumulate=True) for every bin b in one launch, differentiable in ``values``
    (the quantization layer's MLP output) and in ``vox``."""

    @staticmethod
    def forward(ctx, vox, idx, values, bin_stride):
        n, bins = idx.numel(), values.shape[0]
        bad = torch.zeros(1, dtype=torch.int64, device=vox.device)
        with torch.cuda.device(vox.device):
            _lib.check(_lib.load().v2v_put_accumulate_bins(_p(vox), vox.numel(), _p(idx), _p(values), n, bins, int(bin_stride), _p(bad),
                                                           C.c_void_p(torch.cuda.current_stream(vox.device).cuda_stream)))
        ctx.save_for_backward(idx)
        ctx.bin_stride, ctx.bins, ctx.bad = int(bin_stride), bins, bad
        ctx.mark_dirty(vox)
        return vox

    @staticmethod
    def backward(ctx, grad_vox):
        (idx,) = ctx.saved_tensors
        g = grad_vox.contiguous()
        gv = torch.empty((ctx.bins, idx.numel()), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().v2v_take_bins(_p(g), g.numel(), _p(idx), _p(gv), idx.numel(), ctx.bins, ctx.bin_stride,
                                                 C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)))
        return grad_vox, None, gv, None


def put_accumulate_bins(vox: torch.Tensor, idx_before_bins: torch.Tensor, values: torch.Tensor, bin_stride: int) -> torch.Tensor:
    """The scatter of NER-Net's quantization layers: the reference loops over the C temporal bins and calls
    ``vox.put_(clamp(idx_before_bins + W*H*i_bin, max=len(vox)-1), values_i, accumulate=True)`` once per bin
    (model/nernet/representation_modules.py:143-168, 228-248).  Here all bins go in ONE launch and the index tensor is read
    once: ``vox`` flat CUDA float32 (updated in place and returned), ``idx_before_bins`` int64 ``[n]``, ``values`` float32
    ``[C, n]`` (row i = the layer's ``t * value_layer(t - i/(C-1))``), ``bin_stride`` = ``W*H``.  Differentiable in ``values``
    and ``vox`` (the backward pass is the matching gather).  Indices are clamped from above like the reference's; a negative
    index raises ``IndexError`` on the next ``check_put_indices`` (``put_`` would raise)."""
    if not (vox.is_cuda and vox.dtype == torch.float32 and vox.is_contiguous() and vox.dim() == 1):
        raise _lib.V2VError(-1, "vox must be a flat contiguous CUDA float32 tensor (no CPU fallback)")
    idx = idx_before_bins.to(device=vox.device, dtype=torch.int64).contiguous().reshape(-1)
    values = values.to(device=vox.device, dtype=torch.float32)
    if values.dim() == 1:
        values = values[None]
    if values.shape[1] != idx.numel():
        raise ValueError("values must be [C, n] for n indices")
    return _PutAccumulateBins.apply(vox, idx, values.contiguous(), int(bin_stride))


def put_accumulate(vox: torch.Tensor, idx: torch.Tensor, values: torch.Tensor) -> torch.Tensor:
    """``vox.put_(idx, values, accumulate=True)`` for one bin (same kernel, C = 1)."""
    return put_accumulate_bins(vox, idx, values.reshape(1, -1), 0)
