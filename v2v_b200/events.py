"""Event stream -> voxel / image on B200: host side of the scatter path.

Reference-signature functions (same names and argument meaning):

* ``make_voxel``                        data/testh5.py:60-90 (TestH5Dataset.make_voxel)
* ``events_to_voxel_torch``             utils/event_utils.py:466-507
* ``events_to_neg_pos_voxel_torch``     utils/event_utils.py:509-541
* ``events_to_image_torch``             utils/event_utils.py:330-376
* ``events_to_image``                   utils/event_utils.py:155-174
* ``events_to_voxel`` (numpy)           utils/event_utils.py:692-728

plus the batched API ``voxelize_windows`` (one launch for a whole sequence of
frame windows) that the benchmark and a TestH5Dataset drop-in use.  Everything
runs through the C ABI (include/v2v_b200.h); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

_TORCH_DT = {
    torch.uint8: _lib.U8, torch.int8: _lib.I8, torch.int16: _lib.I16, torch.int32: _lib.I32,
    torch.int64: _lib.I64, torch.float32: _lib.F32, torch.float64: _lib.F64, torch.bool: _lib.U8,
}
if hasattr(torch, "uint16"):
    _TORCH_DT[torch.uint16] = _lib.U16


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _to_dev(a, device):
    """numpy / tensor -> contiguous device tensor, keeping the storage dtype."""
    if isinstance(a, torch.Tensor):
        return a.to(device).contiguous()
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint16 and not hasattr(torch, "uint16"):
        a = a.astype(np.int32)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    return torch.from_numpy(a).to(device)


def _dt(t: torch.Tensor) -> int:
    try:
        return _TORCH_DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported event dtype {t.dtype}")


def _env_int(name: str) -> int:
    v = os.environ.get(name)
    return int(v) if v else 0


def _env_flags() -> int:
    """Tuning / test knobs, read here in the Python host per call (the C library reads no environment)."""
    f = 0
    if os.environ.get("V2V_SCATTER_RANGES"):
        f |= _lib.SCATTER_FLAG_RANGES
    if os.environ.get("V2V_SCATTER_GENERIC"):
        f |= _lib.SCATTER_FLAG_GENERIC_SCAN
    if os.environ.get("V2V_SCATTER_PACKED16") == "0":
        f |= _lib.SCATTER_FLAG_NO_PACKED16
    return f


def voxelize_windows(xs, ys, ts, ps, window_offsets, num_bins: int, height: int, width: int, *,
                     mode: str = "h5_discrete", polarity: str = "signed", out_dtype=torch.float32,
                     device="cuda", out: Optional[torch.Tensor] = None, return_dropped: bool = False,
                     validate: Optional[bool] = None, stream: Optional[torch.cuda.Stream] = None, kernel_flags: int = 0):
    """Scatter ``Wn`` windows of one event stream into ``[Wn,bins,H,W]`` in one launch.

    window_offsets: ``[Wn+1]`` ascending event indices (e.g. the ``event_idx``
    attributes of consecutive images, data/testh5.py:113-118).  mode:
    "h5_discrete" | "h5_interp" (data/testh5.py:70-80; ts in seconds, float64
    or float32, ps in {0,1}) | "torch_discrete" | "torch_bilinear"
    (utils/event_utils.py:490-505; float32 arithmetic, ps = signed weights).

    Event order: "h5_interp" takes any order (every event's bin comes from its own timestamp, as in the reference's
    ``np.add.at``).  The other modes assign every bin a contiguous range of the window and need timestamps that are
    non-decreasing inside every window (true for every h5 file the reference's converters write); the library counts
    violations on the device in the same call.  ``validate`` (default: on, unless ``return_dropped`` hands the counters
    back for the caller to check without a sync here) reads that counter and raises ``ValueError`` for unsorted input
    instead of returning mis-binned voxels.  With ``return_dropped`` the result is ``(out, dropped, unsorted)``;
    ``validate=False`` without ``return_dropped`` skips the check pass (one read of the timestamps) altogether.
    """
    dev = torch.device(device)
    modes = {"h5_discrete": _lib.SCATTER_H5_DISCRETE, "h5_interp": _lib.SCATTER_H5_INTERP,
             "torch_discrete": _lib.SCATTER_TORCH_DISCRETE, "torch_bilinear": _lib.SCATTER_TORCH_BILINEAR}
    pols = {"signed": _lib.POL_SIGNED, "pos": _lib.POL_POS_ONLY, "neg": _lib.POL_NEG_ONLY, "split": _lib.POL_SPLIT}
    xs_t, ys_t, ts_t, ps_t = (_to_dev(a, dev) for a in (xs, ys, ts, ps))
    ne = xs_t.numel()
    if not (ys_t.numel() == ne and ts_t.numel() == ne and ps_t.numel() == ne):
        raise AssertionError("xs, ys, ts, ps must have the same length")      # utils/event_utils.py:487
    if mode.startswith("torch") and ts_t.dtype != torch.float32:
        ts_t = ts_t.to(torch.float32)
    if ps_t.dtype not in (torch.uint8, torch.int8, torch.float32):
        ps_t = ps_t.to(torch.float32)
    off_t = _to_dev(np.asarray(window_offsets, dtype=np.int64) if not isinstance(window_offsets, torch.Tensor)
                    else window_offsets.to(torch.int64), dev)
    wn = off_t.numel() - 1
    if wn < 0:
        raise ValueError("window_offsets needs at least one entry")
    if validate is None:
        validate = not return_dropped
    oshape = (wn, 2, num_bins, height, width) if polarity == "split" else (wn, num_bins, height, width)
    if out is None:
        out = torch.empty(oshape, dtype=out_dtype, device=dev)
    elif tuple(out.shape) != oshape or not out.is_contiguous() or not out.is_cuda:
        raise ValueError("out has the wrong shape / layout")
    counters = torch.zeros(2, dtype=torch.int64, device=dev)          # [dropped, unsorted]
    dropped, unsorted = counters[0:1], counters[1:2]

    d = _lib.ScatterDesc()
    d.num_events, d.num_windows = ne, wn
    d.num_bins, d.H, d.W = num_bins, height, width
    d.mode, d.polarity_mode = modes[mode], pols[polarity]
    d.xs_dtype, d.ys_dtype, d.ts_dtype, d.ps_dtype = _dt(xs_t), _dt(ys_t), _dt(ts_t), _dt(ps_t)
    d.out_dtype = {torch.float32: _lib.F32, torch.float64: _lib.F64}[out.dtype]
    d.xs, d.ys, d.ts, d.ps = _ptr(xs_t), _ptr(ys_t), _ptr(ts_t), _ptr(ps_t)
    d.window_offsets, d.voxel, d.dropped = _ptr(off_t), _ptr(out), _ptr(dropped)
    d.unsorted = _ptr(unsorted) if (validate or return_dropped) else None     # validate=False without the counters: no check pass
    d.kernel_flags, d.tuning_splits, d.tuning_smem_kb = int(kernel_flags) | _env_flags(), _env_int("V2V_SCATTER_SPLITS"), _env_int("V2V_SCATTER_SMEM_KB")
    work = torch.empty((int(_lib.load().v2v_scatter_workspace_bytes(C.byref(d))) + 15) // 16 * 2, dtype=torch.int64, device=dev)
    d.workspace, d.workspace_bytes = _ptr(work), work.numel() * 8
    s = stream if stream is not None else torch.cuda.current_stream(dev)
    if stream is not None:       # conversions, counters and H2D copies above were enqueued on the current stream
        stream.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_events_to_voxel(C.byref(d), C.c_void_p(s.cuda_stream)))
    for t in (xs_t, ys_t, ts_t, ps_t, off_t, work, counters):
        t.record_stream(s)
    if validate and int(unsorted.item()) != 0:
        raise ValueError(f"timestamps decrease at {int(unsorted.item())} positions inside windows: mode {mode!r} needs "
                         "non-decreasing timestamps inside every window")
    if return_dropped:
        return out, dropped, unsorted
    return out


def make_voxel(evs: Sequence, num_bins: int, height: int, width: int, interpolate_bins: bool = False,
               device="cuda") -> np.ndarray:
    """``TestH5Dataset.make_voxel`` (data/testh5.py:60-90): ``evs = [ts, xs, ys, ps]`` of
    one window in the h5 dtypes -> float64 ``[bins,H,W]``."""
    ts, xs, ys, ps = evs
    n = int(np.asarray(ts).shape[0]) if not isinstance(ts, torch.Tensor) else ts.numel()
    if n == 0:                                                             # :63-64
        return np.zeros((num_bins, height, width))
    out = voxelize_windows(xs, ys, ts, ps, [0, n], num_bins, height, width,
                           mode="h5_interp" if interpolate_bins else "h5_discrete",
                           out_dtype=torch.float64, device=device)
    return out[0].cpu().numpy()


class MakeVoxelMixin:
    """Override of ``make_voxel`` for the reference's TestH5 datasets.

    ``class TestH5DatasetB200(MakeVoxelMixin, TestH5Dataset): pass`` — the
    dataset keeps reading h5 on the CPU; only the scatter moves to the GPU.
    Uses ``self.num_bins, self.H, self.W, self.interpolate_bins``
    (data/testh5.py:29-39).
    """

    v2v_device = "cuda"

    def make_voxel(self, evs):
        return make_voxel(evs, self.num_bins, self.H, self.W, self.interpolate_bins, device=self.v2v_device)


def events_to_voxel_torch(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240), temporal_bilinear=True):
    """Same contract as utils/event_utils.py:466-507; returns a float32 ``[B,H,W]``
    tensor on ``device`` (default: CUDA — the reference allocates on the CPU
    regardless of ``device``)."""
    if device is None:
        device = xs.device if isinstance(xs, torch.Tensor) and xs.is_cuda else "cuda"
    assert len(xs) == len(ys) and len(ys) == len(ts) and len(ts) == len(ps)
    n = len(xs)
    out = voxelize_windows(xs, ys, ts, ps, [0, n], B, sensor_size[0], sensor_size[1],
                           mode="torch_bilinear" if temporal_bilinear else "torch_discrete", device=device)
    return out[0]


def events_to_neg_pos_voxel_torch(xs, ys, ts, ps, B, device=None, sensor_size=(180, 240), temporal_bilinear=True):
    """utils/event_utils.py:509-541: (voxel_pos, voxel_neg) with 0/1 weights."""
    if device is None:
        device = xs.device if isinstance(xs, torch.Tensor) and xs.is_cuda else "cuda"
    n = len(xs)
    mode = "torch_bilinear" if temporal_bilinear else "torch_discrete"
    both = voxelize_windows(xs, ys, ts, ps, [0, n], B, sensor_size[0], sensor_size[1], mode=mode, polarity="split",
                            device=device)             # one launch: [1, 2, B, H, W]
    return both[0, 0], both[0, 1]


def _image(xs, ys, ps, sensor_size, bilinear, padding, clip, out_dtype, device):
    dev = torch.device(device)
    xs_t, ys_t = _to_dev(xs, dev), _to_dev(ys, dev)
    ps_t = None if ps is None else _to_dev(ps, dev)
    if ps_t is not None:
        ps_t = ps_t.reshape(-1)
        if ps_t.dtype not in (torch.float32, torch.float64):
            ps_t = ps_t.to(torch.float32)
        if out_dtype != torch.float64 and ps_t.dtype == torch.float64:
            ps_t = ps_t.to(torch.float32)
    h, w = sensor_size
    ho, wo = (h + 1, w + 1) if (bilinear and padding) else (h, w)
    img = torch.empty((ho, wo), dtype=out_dtype, device=dev)
    dropped = torch.zeros(1, dtype=torch.int64, device=dev)
    d = _lib.ImageDesc()
    d.num_events, d.H, d.W = xs_t.numel(), h, w
    d.bilinear, d.padding, d.clip_out_of_range = int(bilinear), int(padding), int(clip)
    d.xs_dtype, d.ys_dtype = _dt(xs_t), _dt(ys_t)
    d.ps_dtype = _dt(ps_t) if ps_t is not None else _lib.F32
    d.out_dtype = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.int64: _lib.I64}[out_dtype]
    d.xs, d.ys, d.ps, d.image, d.dropped = _ptr(xs_t), _ptr(ys_t), _ptr(ps_t), _ptr(img), _ptr(dropped)
    s = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_events_to_image(C.byref(d), C.c_void_p(s.cuda_stream)))
    return img


def events_to_image_torch(xs, ys, ps, device=None, sensor_size=(180, 240), clip_out_of_range=True,
                          interpolation=None, padding=True):
    """utils/event_utils.py:330-376.  Nearest: coordinates truncated, the clip
    mask is not applied (as in the reference); bilinear: 4-tap splat into an
    image padded by one row/column when ``padding``."""
    if device is None:
        device = xs.device if isinstance(xs, torch.Tensor) and xs.is_cuda else "cuda"
    is_float = (xs.dtype.is_floating_point if isinstance(xs, torch.Tensor) else np.asarray(xs).dtype.kind == "f")
    bilinear = interpolation == "bilinear" and is_float
    if interpolation == "bilinear" and padding and not bilinear:
        # integer coordinates with bilinear requested: the reference still allocates the padded image
        img = _image(xs, ys, ps, (sensor_size[0] + 1, sensor_size[1] + 1), False, False, False, torch.float32, device)
        return img
    return _image(xs, ys, ps, sensor_size, bilinear, padding, clip_out_of_range, torch.float32, device)


def events_to_image(xs, ys, ps, sensor_size=(180, 240), interpolation=None, padding=False, device="cuda"):
    """NumPy flavour (utils/event_utils.py:155-174): float64 image via nearest-pixel accumulation."""
    if interpolation == "bilinear" and np.asarray(xs).dtype.kind == "f":
        return events_to_image_torch(torch.from_numpy(np.asarray(xs, dtype=np.float32)),
                                     torch.from_numpy(np.asarray(ys, dtype=np.float32)),
                                     torch.from_numpy(np.asarray(ps, dtype=np.float32)), device=device,
                                     sensor_size=sensor_size, clip_out_of_range=True, interpolation="bilinear",
                                     padding=padding).cpu().numpy()
    img = _image(xs, ys, np.asarray(ps, dtype=np.float64), sensor_size, False, False, False, torch.float64, device)
    return img.cpu().numpy()


def events_to_voxel(xs, ys, ts, ps, B, sensor_size=(180, 240), temporal_bilinear=True, device="cuda"):
    """NumPy flavour (utils/event_utils.py:692-728), bilinear branch only: the reference's
    non-bilinear branch reads ``weights`` before assignment.  float64 in/out."""
    if not temporal_bilinear:
        raise NotImplementedError("the reference's non-bilinear numpy branch raises UnboundLocalError")
    ts = np.asarray(ts, dtype=np.float64).reshape(-1)
    ps = np.asarray(ps, dtype=np.float64).reshape(-1)
    span = ts[-1] - ts[0]
    tn = (ts - ts[0]) / span * (B - 1)
    planes = []
    for bi in range(B):
        wgt = ps * np.maximum(0.0, 1.0 - np.abs(tn - bi))
        planes.append(_image(xs, ys, wgt, sensor_size, False, False, False, torch.float64, device))
    return torch.stack(planes).cpu().numpy()


def event_count_map(xs, ys, height, width, device="cuda") -> torch.Tensor:
    """Per-pixel int64 event counts (scripts/testset_evcnt_maps.py:19-25)."""
    return _image(xs, ys, None, (height, width), False, False, False, torch.int64, device)


def fps_window_offsets(ts, fps: float, device="cuda"):
    """Event offsets of fixed-rate windows, as ``FPS_H5Dataset.__init__`` computes them (data/testh5.py:468-474):
    ``total = int((ts[-1]-ts[0])*FPS)``, ``borders = np.linspace(ts[0], ts[-1], total+1)``,
    ``event_idx = np.searchsorted(ts, borders)``.  ``ts`` float64 seconds (sorted), host or device.
    Returns (event_idx int64 [total+1] on the device, borders float64 numpy)."""
    dev = torch.device(device)
    ts_t = _to_dev(ts, dev)
    if ts_t.dtype != torch.float64:
        ts_t = ts_t.to(torch.float64)
    n = ts_t.numel()
    if n == 0:
        return torch.zeros(1, dtype=torch.int64, device=dev), np.zeros(1)
    lo_hi = ts_t[[0, n - 1]].cpu().numpy()
    total = int((lo_hi[1] - lo_hi[0]) * fps)
    borders = np.linspace(lo_hi[0], lo_hi[1], total + 1)            # the 2-number host step keeps numpy's exact linspace
    b_t = torch.from_numpy(borders).to(dev)
    out = torch.empty(total + 1, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_searchsorted_f64(_ptr(ts_t), n, _ptr(b_t), total + 1, _ptr(out), C.c_void_p(s.cuda_stream)))
    return out, borders


def pack_events_n5(xs, ys, ts, ps, device="cuda") -> torch.Tensor:
    """Raw event tensor of the NER-Net loader (data/testh5.py:329-339): float64 ``[N,5]`` = [x, y, t, 2p-1, 0];
    an empty window gives ``zeros((1,5))`` like the reference (:341-342)."""
    dev = torch.device(device)
    xs_t, ys_t, ts_t, ps_t = (_to_dev(a, dev) for a in (xs, ys, ts, ps))
    n = xs_t.numel()
    if n == 0:
        return torch.zeros((1, 5), dtype=torch.float64, device=dev)
    out = torch.empty((n, 5), dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().v2v_pack_events_n5(_ptr(xs_t), _dt(xs_t), _ptr(ys_t), _dt(ys_t), _ptr(ts_t), _dt(ts_t),
                                                  _ptr(ps_t), _dt(ps_t), n, _ptr(out), C.c_void_p(s.cuda_stream)))
    return out
