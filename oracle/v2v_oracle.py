"""CPU oracle for the V2V video-to-voxel hot path (TEST INFRASTRUCTURE ONLY).

This module is a plain-NumPy restatement of the reference's algorithms for the
hot path.  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package ``v2v_b200`` never imports anything under ``oracle/``.

Parity status: the reference ships no tests or golden vectors, so
parity is pinned against outputs of the *unmodified reference run in the build
container*: ``tests/golden/make_golden.py`` imports ``/root/reference`` with the
stubs of SURVEY.md Appendix B, asserts that every function below is
bit-identical to the reference under identical ``np.random.seed`` streams, and
writes the golden ``.npz`` fixtures that the GPU-side tests replay.

All randomness is an *explicit input* here (the reference draws from the global
legacy ``np.random`` stream); ``esim_draw_randomness`` and the ``record`` hook of
``v2e_video_to_voxel`` reproduce the reference's draw order on a given ``RandomState``-like object so
that "same seed → same result" can be checked bit for bit.

Citations are relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# ESIM-style frame -> voxel (data/v2v_core_esim.py)
# --------------------------------------------------------------------------


def esim_log_lut() -> np.ndarray:
    """256-entry float64 table of F2∘F1: log(0.001 + ((v/255)**2.2*255)/255).

    Follows data/v2v_core_esim.py:3-4 (reverse_gamma_correction) and :34 (log).
    Built with the same NumPy array expressions the reference applies to the
    whole clip, so ``lut[video]`` is bit-identical to the reference's
    ``log_imgs`` on the same host (the values are host/SIMD dependent at the
    1-ulp level, which is why every consumer takes the LUT as an input).
    """
    v = np.arange(256, dtype=np.uint8)
    lin = (v / 255) ** 2.2 * 255
    return np.log(0.001 + lin / 255.0)


def esim_draw_randomness(n_frames, height, width, hot_pixel_fraction, hot_pixel_std, rs=np.random):
    """Draw (U0, hot_noise, G) in the reference's order from ``rs``.

    Order on the stream (data/v2v_core_esim.py:29,37,38,44): rand(H,W) for the
    initial potential, rand(H,W) for the hot-pixel mask, randn(H,W) for the hot
    pixel amplitudes, then one randn(H,W) per interval.  Drawing the per-interval
    fields in one call yields the same stream as N-1 successive calls.
    """
    u0 = rs.rand(height, width)
    mask = rs.rand(height, width) < hot_pixel_fraction
    hot = hot_pixel_std * rs.randn(height, width)
    hot = np.where(mask, hot, 0)
    g = rs.randn(n_frames - 1, height, width)
    return u0, hot, g


def esim_video_to_voxel(video, pos_thres, neg_thres, base_noise_std, u0, hot_noise, g,
                        put_noise_external=False, lut=None, return_state=False):
    """Per-interval signed crossing counts, float64 [N-1,H,W].

    Restates EventEmulator.video_to_voxel (data/v2v_core_esim.py:26-69) with the
    random fields passed in: ``u0`` uniform [0,1) [H,W]; ``hot_noise`` [H,W]
    (already masked and scaled, :37-39); ``g`` standard normal [N-1,H,W].
    """
    video = np.asarray(video)
    assert video.dtype == np.uint8 and video.ndim == 3
    n = video.shape[0]
    if lut is None:
        lut = esim_log_lut()
    logv = lut[video]                                         # == reference log_imgs
    pot = u0 * (pos_thres + neg_thres) - neg_thres            # :29
    out = np.empty((n - 1,) + video.shape[1:], dtype=np.float64)
    for i in range(n - 1):
        pot = pot + (logv[i + 1] - logv[i])                   # :42-43
        bn = base_noise_std * g[i]                            # :44
        if not put_noise_external:                            # :46-49
            pot = pot + bn
            pot = pot + hot_noise
        pe = np.where(pot >= pos_thres, np.floor_divide(pot, pos_thres), 0)      # :51-52
        ne = np.where(pot <= -neg_thres, np.floor_divide(-pot, neg_thres), 0)    # :54-55
        pot = pot - pe * pos_thres                            # :57
        pot = pot + ne * neg_thres                            # :58
        vox = pe - ne                                         # :60
        if put_noise_external:                                # :62-65
            vox = vox + bn
            vox = vox + hot_noise
        out[i] = vox
    if return_state:
        return out, pot
    return out


def bin_accumulate(intervals, num_bins, frames_per_bin):
    """[N-1,H,W] -> [T,bins,H,W]: sum ``frames_per_bin`` consecutive intervals.

    data/v2v_datasets.py:365-366,399-400.
    """
    n1, h, w = intervals.shape
    group = num_bins * frames_per_bin
    if n1 % group != 0:
        raise AssertionError("(N-1) must be a multiple of num_bins*frames_per_bin")
    t = n1 // group
    return intervals.reshape(t, num_bins, frames_per_bin, h, w).sum(axis=2)


def sample_esim_params(rs=np.random, threshold_range=(0.05, 2), max_thres_pos_neg_gap=1.5,
                       base_noise_std_range=(0, 0.2), hot_pixel_fraction_range=(0, 0.001),
                       hot_pixel_std_range=(0, 0.2), scale_noise_strength=False,
                       put_noise_external=False, pos_thres=None, neg_thres=None):
    """Sampling law of the simulator parameters (data/v2v_datasets.py:368-386).

    When ``pos_thres``/``neg_thres`` are given (use_fixed_thresholds) no
    threshold draws are consumed.
    """
    if pos_thres is None or neg_thres is None:
        a = rs.uniform(*threshold_range)
        gap = rs.uniform(1, max_thres_pos_neg_gap)
        b = a * gap
        if rs.rand() > 0.5:
            pos_thres, neg_thres = a, b
        else:
            pos_thres, neg_thres = b, a
    base_noise_std = rs.uniform(*base_noise_std_range)
    hot_pixel_fraction = rs.uniform(*hot_pixel_fraction_range)
    hot_pixel_std = rs.uniform(*hot_pixel_std_range)
    if scale_noise_strength and not put_noise_external:
        base_noise_std = base_noise_std * pos_thres
        hot_pixel_std = hot_pixel_std * pos_thres
    return {
        "pos_thres": pos_thres,
        "neg_thres": neg_thres,
        "base_noise_std": base_noise_std,
        "hot_pixel_fraction": hot_pixel_fraction,
        "hot_pixel_std": hot_pixel_std,
    }


def pack_frames(all_imgs, frames_per_img, img_cnt, output_additional_frame=False):
    """Ground-truth frame tensor of a training sample, float32 [T(+1),C,H,W] in [0,1].

    data/v2v_datasets.py:328-338,352: every ``frames_per_img``-th image (starting
    at index ``frames_per_img``, or at 0 with output_additional_frame), HWC->CHW,
    float32, divided by 255 in float32.
    """
    if output_additional_frame:
        idx = [i * frames_per_img for i in range(img_cnt + 1)]
    else:
        idx = [(i + 1) * frames_per_img for i in range(img_cnt)]
    sel = np.stack([all_imgs[i] for i in idx]).astype(np.float32)      # [T,H,W,C]
    sel = np.transpose(sel, (0, 3, 1, 2))
    return sel / np.float32(255)


# --------------------------------------------------------------------------
# v2e-style frame -> voxel (data/v2v_core_v2e.py)
# --------------------------------------------------------------------------


def v2e_log_lut() -> np.ndarray:
    """float32 table of the effective lin_log: float32(log(v/255 + 0.01)).

    data/v2v_core_v2e.py:120-137 — the piecewise lin/log result is overwritten by
    the plain log at :135, evaluated in float64 and returned as float32.
    """
    v = np.arange(256, dtype=np.float64)
    return np.log(v / 255 + 0.01).astype(np.float32)


def _v2e_thresholds(params, a, b):
    """Per-pixel (pos, neg) maps and their shot-noise pre-probabilities.

    data/v2v_core_v2e.py:333-343 (+ clip and nominal/thres at :392-399).
    """
    if params["threshold_model"] == "pn_related":
        pos = a + (b / 2)
        neg = a - (b / 2)
    else:
        pos, neg = a, b
    pos = np.clip(pos, a_min=0.01, a_max=None)
    neg = np.clip(neg, a_min=0.01, a_max=None)
    pos_nom = params["thres_mean_mean"] + params["thres_diff_mean"] / 2     # :297
    neg_nom = params["thres_mean_mean"] - params["thres_diff_mean"] / 2     # :298
    return pos, neg, np.divide(pos_nom, pos), np.divide(neg_nom, neg)


def v2e_video_to_voxel(video, fps, params, rs=np.random, lut=None, record=None, maps=None):
    """v2e-style simulation, float64 [N-1,H,W] of (pos - neg) counts per interval.

    Restates video_to_voxel / EventEmulator.generate_events
    (data/v2v_core_v2e.py:401-581).  ``video`` is [N,H,W] with integer values
    0..255; a uint8 array takes the reference's uint8 arithmetic in
    ``rescale_intensity_frame`` (:190: ``new_frame+20`` wraps for values >= 236), any other dtype is
    converted to float64 first (no wrap).  ``params`` keys: threshold_model, thres_mean_mean,
    thres_mean_std, thres_diff_mean, thres_diff_std, cutoff_hz, leak_rate_hz,
    shot_noise_rate_hz, leak_jitter_fraction, noise_rate_cov_decades
    (refractory_period_s must be 0: the reference's branch raises TypeError).

    Randomness is drawn from ``rs`` in the reference's order; if ``record`` is a
    dict every drawn field is stored in it (lists per frame) so that a GPU run
    can replay exactly the same fields:  thr_a, thr_b, noise_randn, and per
    frame k>=1: leak_randn[k-1], pos_shot[k-1], neg_shot[k-1].

    ``maps=(pos_thres, neg_thres, noise_rate)`` replaces the first-frame draws
    (:333-349) by given per-pixel maps (already clipped) — used to replay GPU
    runs whose maps were prepared elsewhere; no draw is consumed for them.
    """
    video = np.asarray(video)
    n, h, w = video.shape
    shape = (h, w)
    model = params["threshold_model"]
    if model not in ("pn_related", "spatial_independent", "spatial_temporal_independent"):
        raise NotImplementedError("spatial_independent_temporal_changing: see DESIGN.md")
    per_frame = model == "spatial_temporal_independent"       # thresholds re-drawn on every frame (:417-421)
    cutoff = params["cutoff_hz"]
    leak_hz = params["leak_rate_hz"]
    shot_hz = params["shot_noise_rate_hz"]
    jitter = params["leak_jitter_fraction"]
    if lut is None:
        lut = v2e_log_lut()
    if record is not None:
        record.update({"leak_randn": [], "pos_shot": [], "neg_shot": [], "pos_thres_frames": [], "neg_thres_frames": []})
    vid_idx = video.astype(np.int64)
    out = np.empty((n - 1, h, w), dtype=np.float64)
    t_prev = 0.0
    lp = base = None
    for k in range(n):
        t_k = k / fps
        dt = t_k - t_prev                                             # :442
        if per_frame and maps is None:                                # :417-421, before anything else, on EVERY frame
            a = rs.normal(loc=params["thres_mean_mean"], scale=params["thres_mean_std"], size=shape)
            b = rs.normal(loc=params["thres_mean_mean"], scale=params["thres_mean_std"], size=shape)
            pos_thr, neg_thr, pos_pp, neg_pp = _v2e_thresholds(params, a, b)
            if record is not None and k > 0:
                record["pos_thres_frames"].append(pos_thr)
                record["neg_thres_frames"].append(neg_thr)
        frame = video[k].astype(np.float64)
        log_new = lut[vid_idx[k]]                                     # float32, :447
        inten01 = None
        if cutoff > 0 or shot_hz > 0:
            inten01 = (((vid_idx[k] + 20) & 255) if video.dtype == np.uint8 else (frame + 20)) / 275.   # :455,190
        if base is None:
            lp = log_new                                              # :463-465
        if cutoff > 0:                                                # :157-173
            tau = 1 / (math.pi * 2 * cutoff)
            eps = inten01 * (dt / tau)
            eps = np.clip(eps, a_min=None, a_max=1)
            lp = (1 - eps) * lp + eps * log_new
        else:
            lp = log_new
        if base is None and maps is not None:
            pos_thr, neg_thr, noise_rate = maps
            pos_nom = params["thres_mean_mean"] + params["thres_diff_mean"] / 2
            neg_nom = params["thres_mean_mean"] - params["thres_diff_mean"] / 2
            pos_pp, neg_pp = np.divide(pos_nom, pos_thr), np.divide(neg_nom, neg_thr)
            base = lp
            continue
        if base is None:                                              # :474-478
            a = rs.normal(loc=params["thres_mean_mean"], scale=params["thres_mean_std"], size=shape)
            if model == "pn_related":
                b = rs.normal(loc=params["thres_diff_mean"], scale=params["thres_diff_std"], size=shape)
            else:
                b = rs.normal(loc=params["thres_mean_mean"], scale=params["thres_mean_std"], size=shape)
            pos_thr, neg_thr, pos_pp, neg_pp = _v2e_thresholds(params, a, b)      # (per-frame model: replaced on the next frame)
            nr = rs.randn(*shape).astype(np.float32)                  # :348
            noise_rate = np.exp(math.log(10) * params["noise_rate_cov_decades"] * nr)   # :349 (float32)
            if record is not None:
                record.update({"thr_a": a, "thr_b": b, "noise_randn": nr,
                               "pos_thres": pos_thr, "neg_thres": neg_thr, "noise_rate": noise_rate})
            base = lp
            continue
        if leak_hz > 0:                                               # :487-494,192-211
            r = rs.randn(h, w)
            rate = leak_hz * noise_rate * (1 - jitter * r)
            base = base - dt * rate * pos_thr
            if record is not None:
                record["leak_randn"].append(r)
        diff = lp - base                                              # :503
        pos_n = np.floor_divide(np.clip(diff, a_min=0, a_max=None), pos_thr)    # :55-60
        neg_n = np.floor_divide(np.clip(-diff, a_min=0, a_max=None), neg_thr)
        if shot_hz > 0:                                               # :65-105,517-525
            inten_factor = 1 - (1 - 0.25) * inten01
            pf = inten_factor * pos_pp
            pf = pf / np.mean(pf)
            nf = inten_factor * neg_pp
            nf = nf / np.mean(nf)
            sf = (shot_hz / 2) * dt
            ps_n = rs.poisson(pf * sf)
            ns_n = rs.poisson(nf * sf)
            if record is not None:
                record["pos_shot"].append(ps_n)
                record["neg_shot"].append(ns_n)
            pos_n = pos_n + ps_n
            neg_n = neg_n + ns_n
        # in-place semantics of :547-548: result is cast back to base's dtype
        base = (base + pos_n * pos_thr).astype(base.dtype)
        base = (base - neg_n * neg_thr).astype(base.dtype)
        t_prev = t_k                                                  # :551
        out[k - 1] = pos_n - neg_n                                    # :579-580
    return out


def v2e_replay(video, fps, params, fields, lut=None):
    """Same as ``v2e_video_to_voxel`` but replaying recorded random fields.

    ``fields`` is the ``record`` dict of a previous run (thr_a, thr_b,
    noise_randn, leak_randn[], pos_shot[], neg_shot[]).
    """

    class _Replay:
        def __init__(self, f):
            self.f = f
            self.normals = [f["thr_a"], f["thr_b"]]
            self.leak = list(f.get("leak_randn", []))
            self.shots = []
            for p, q in zip(f.get("pos_shot", []), f.get("neg_shot", [])):
                self.shots += [p, q]
            self.first_randn = True

        def normal(self, loc, scale, size):
            return self.normals.pop(0)

        def randn(self, *shape):
            if self.first_randn:
                self.first_randn = False
                return self.f["noise_randn"].astype(np.float64)
            return self.leak.pop(0)

        def poisson(self, lam):
            return self.shots.pop(0)

    return v2e_video_to_voxel(video, fps, params, rs=_Replay(fields), lut=lut)


# --------------------------------------------------------------------------
# Event stream -> voxel: test-loop flavour (data/testh5.py:60-90)
# --------------------------------------------------------------------------


def make_voxel(ts, xs, ys, ps, num_bins, height, width, interpolate_bins=False):
    """float64 [bins,H,W] voxel of one event window (TestH5Dataset.make_voxel).

    ``ts`` seconds in the h5 dtype (float64, or float32 for EVAID), ``xs``/``ys``
    integer pixel coordinates, ``ps`` in {0,1}.  data/testh5.py:60-90.
    """
    vox = np.zeros((num_bins, height, width))
    ts = np.asarray(ts)
    if ts.shape[0] == 0:                                              # :63-64
        return vox
    pol = np.asarray(ps).astype(np.int8) * 2 - 1                      # :67
    tau = ((ts - ts[0]) * 1e6).astype(np.int64)                       # :68 (µs, truncation)
    ys = np.asarray(ys)
    xs = np.asarray(xs)
    if not interpolate_bins:
        width_us = (tau[-1] + 0.001) / num_bins                       # :71
        b = np.floor(tau / width_us).astype(np.uint8)                 # :72
        np.add.at(vox, (b, ys, xs), pol)                              # :73
    else:
        span = tau[-1] - tau[0]                                       # :76
        tn = (tau - tau[0]) / (span + 0.0001) * (num_bins - 1)        # :77
        for bi in range(num_bins):                                    # :78-80
            wgt = np.maximum(0, 1.0 - np.abs(tn - bi))
            np.add.at(vox, (bi, ys, xs), wgt * pol)
    return vox


# --------------------------------------------------------------------------
# Event stream -> voxel / image: legacy torch flavour (utils/event_utils.py)
# All arithmetic is float32, exactly as torch CPU performs it.
# --------------------------------------------------------------------------

_F = np.float32


def events_to_image_f32(xs, ys, ps, sensor_size=(180, 240), clip_out_of_range=True,
                        interpolation=None, padding=True):
    """float32 event image (utils/event_utils.py:330-376, 176-184).

    ``xs``/``ys`` float32 (or integer) coordinates, ``ps`` float32 weights.
    Nearest: truncate coordinates, sequential float32 accumulation.  Bilinear:
    4-tap spatial splat into an image padded by one row/column when ``padding``.
    Out-of-range events are handled like the reference: in the bilinear branch
    coordinates and weights are multiplied by a 0/1 mask (so they land on pixel
    (0,0) with weight 0); the nearest branch computes the mask but never applies
    it (:371-375), so out-of-range events are an IndexError there as in torch.
    """
    xs = np.asarray(xs)
    ys = np.asarray(ys)
    ps = np.asarray(ps, dtype=_F).reshape(-1)
    bilinear = interpolation == "bilinear" and xs.dtype.kind == "f"
    if interpolation == "bilinear" and padding:
        h, w = sensor_size[0] + 1, sensor_size[1] + 1
    else:
        h, w = sensor_size
    mask = np.ones(xs.shape, dtype=_F)
    if clip_out_of_range:                                             # :353-358
        cx = w if (interpolation is None and padding is False) else w - 1
        cy = h if (interpolation is None and padding is False) else h - 1
        mask = np.where(xs >= cx, _F(0), _F(1)) * np.where(ys >= cy, _F(0), _F(1))
    img = np.zeros((h, w), dtype=_F)
    if bilinear:                                                      # :361-369
        xf = xs.astype(_F)
        yf = ys.astype(_F)
        px = np.floor(xf)
        py = np.floor(yf)
        dx = (xf - px).astype(_F)
        dy = (yf - py).astype(_F)
        ix = (px * mask).astype(np.int64)
        iy = (py * mask).astype(np.int64)
        wgt = (ps * mask).astype(_F)
        one = _F(1.0)
        np.add.at(img, (iy, ix), wgt * (one - dx) * (one - dy))
        np.add.at(img, (iy, ix + 1), wgt * dx * (one - dy))
        np.add.at(img, (iy + 1, ix), wgt * (one - dx) * dy)
        np.add.at(img, (iy + 1, ix + 1), wgt * dx * dy)
    else:                                                             # :371-375
        ix = xs.astype(np.int64)
        iy = ys.astype(np.int64)
        np.add.at(img, (iy, ix), ps)
    return img


def events_to_voxel_f32(xs, ys, ts, ps, num_bins, sensor_size=(180, 240), temporal_bilinear=True):
    """float32 [B,H,W] voxel (events_to_voxel_torch, utils/event_utils.py:466-507)."""
    xs = np.asarray(xs)
    ys = np.asarray(ys)
    ts = np.asarray(ts, dtype=_F)
    ps = np.asarray(ps, dtype=_F)
    assert len(xs) == len(ys) == len(ts) == len(ps)
    h, w = sensor_size
    span = _F(ts[-1] - ts[0])                                         # :489
    rel = (ts - ts[0]).astype(_F)
    if temporal_bilinear:
        with np.errstate(divide="ignore", invalid="ignore"):
            tn = (rel / span * _F(num_bins - 1)).astype(_F)           # :490
        planes = []
        for bi in range(num_bins):                                    # :493-499
            wgt = np.maximum(_F(0), _F(1.0) - np.abs(tn - _F(bi))).astype(_F)
            planes.append(events_to_image_f32(xs, ys, ps * wgt, sensor_size=sensor_size,
                                              clip_out_of_range=False))
        return np.stack(planes)
    vox = np.zeros((num_bins, h, w), dtype=_F)                        # :502-505
    per_bin = _F(_F(span + _F(0.001)) / _F(num_bins))
    b = np.floor(rel / per_bin).astype(np.int32)
    np.add.at(vox, (b, ys.astype(np.int32), xs.astype(np.int32)), ps)
    return vox


def events_to_neg_pos_voxel_f32(xs, ys, ts, ps, num_bins, sensor_size=(180, 240), temporal_bilinear=True):
    """(pos, neg) voxels with 0/1 weights (utils/event_utils.py:509-541)."""
    ps = np.asarray(ps, dtype=_F)
    pw = np.where(ps > 0, _F(1), _F(0))
    nw = np.where(ps <= 0, _F(1), _F(0))
    return (events_to_voxel_f32(xs, ys, ts, pw, num_bins, sensor_size, temporal_bilinear),
            events_to_voxel_f32(xs, ys, ts, nw, num_bins, sensor_size, temporal_bilinear))


def events_to_image_np(xs, ys, ps, sensor_size=(180, 240)):
    """NumPy nearest-pixel event image via bincount, float64 (utils/event_utils.py:155-174)."""
    flat = np.ravel_multi_index(np.stack((ys, xs)), sensor_size)
    return np.bincount(flat, weights=ps, minlength=sensor_size[0] * sensor_size[1]).reshape(sensor_size)


def events_to_voxel_np(xs, ys, ts, ps, num_bins, sensor_size=(180, 240)):
    """NumPy float64 voxel with temporal bilinear weights (events_to_voxel, utils/event_utils.py:692-728, the
    ``temporal_bilinear=True`` branch; the other branch reads ``weights`` before assignment in the reference).
    ``ts``/``ps`` may be ``[N]`` or the ``[N,1]`` columns the reference needs."""
    ts = np.asarray(ts, dtype=np.float64).reshape(-1)
    ps = np.asarray(ps, dtype=np.float64).reshape(-1)
    span = ts[-1] - ts[0]                                            # :711
    tn = (ts - ts[0]) / span * (num_bins - 1)                        # :712
    planes = []
    for bi in range(num_bins):                                       # :714-726
        wgt = ps * np.maximum(0.0, 1.0 - np.abs(tn - bi))
        planes.append(events_to_image_np(xs, ys, wgt, sensor_size))
    return np.stack(planes)


def event_count_map(xs, ys, height, width):
    """Per-pixel event count (scripts/testset_evcnt_maps.py:19-25)."""
    cnt = np.zeros((height, width), dtype=np.int64)
    np.add.at(cnt, (np.asarray(ys).astype(np.int64), np.asarray(xs).astype(np.int64)), 1)
    return cnt


# ---- frame-side packing of the dataset ------------------------------------------------------

def pause_indices(count, proba_pause_when_running, proba_pause_when_paused, rs=np.random):
    """Pause sequence of WebvidDatasetV2.__getitem__ (data/v2v_datasets.py:285-301): ``count`` indices into the raw
    clip, one ``rs.rand()`` per frame in the reference's order.  Returns (img_idxes, true_img_cnt)."""
    img_idxes, idx, is_pause = [], 0, False
    for _ in range(count):
        img_idxes.append(idx)
        if is_pause and rs.rand() > proba_pause_when_paused:
            is_pause = False
        elif not is_pause and rs.rand() < proba_pause_when_running:
            is_pause = True
        if not is_pause:
            idx += 1
    return np.asarray(img_idxes, dtype=np.int64), idx + 1


def degrade_video(imgs, kind, rs=np.random):
    """HDR / LDR degrade (data/v2v_datasets.py:473-483): one ``rs.uniform`` for the scale, then per frame
    ``np.clip((img-127.5)*scale+127.5, 0, 255).astype(np.uint8)``.  Returns (list of frames, scale)."""
    scale = rs.uniform(1, 3) if kind == "hdr" else rs.uniform(0.3, 1)
    return [np.clip((im - 127.5) * scale + 127.5, 0, 255).astype(np.uint8) for im in imgs], scale


def bgr_to_gray(img_stack):
    """data/v2v_datasets.py:19-22 (the float64 summation order is NumPy's / the BLAS kernel's: pinned by tests/golden)."""
    return np.dot(img_stack[..., :3], [0.5870, 0.1140, 0.2989]).astype(np.uint8)


# ---- published generators, restated for the known-answer tests of the in-kernel RNG --------------------------------

def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(x) & 0xFFFFFFFF for x in counter]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def xoshiro128pp(state, n):
    """First ``n`` outputs of xoshiro128++ 1.0 (Blackman & Vigna, 2019) from a 4-word state."""
    s = [int(x) & 0xFFFFFFFF for x in state]
    rotl = lambda x, r: ((x << r) | (x >> (32 - r))) & 0xFFFFFFFF
    out = []
    for _ in range(n):
        out.append((rotl((s[0] + s[3]) & 0xFFFFFFFF, 7) + s[0]) & 0xFFFFFFFF)
        t = (s[1] << 9) & 0xFFFFFFFF
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = rotl(s[3], 11)
    return out


ESIM_LCG_MUL = 0xF9B25D65      # 32-bit multiplier for a 64-bit LCG (Steele & Vigna 2021), v2v_b200/csrc/esim_common.cuh


def esim_noise_stream_words(seed, clip_index, pixel_group, n):
    """The ESIM base-noise stream of one (clip, 4-pixel group) (v2v_b200/csrc/esim_common.cuh): the Philox4x32-10 block
    r of counter (group lo32, 0, clip lo32, tag1 | group hi14 << 16 | clip hi16) under key (seed lo32, seed hi32)
    seeds the 64-bit linear congruential generator s' = s*0xf9b25d65 + c (mod 2^64) with s = r[1]<<32 | r[0] and
    the odd increment c = r[3]<<32 | r[2] | 1; the outputs are the high words of the successive states."""
    ctr = [pixel_group & 0xFFFFFFFF, 0, clip_index & 0xFFFFFFFF,
           0x40000000 | (((pixel_group >> 32) & 0x3FFF) << 16) | ((clip_index >> 32) & 0xFFFF)]
    r = philox4x32_10(ctr, [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF])
    s, c = (r[1] << 32) | r[0], (r[3] << 32) | r[2] | 1
    out = []
    for _ in range(n):
        s = (s * ESIM_LCG_MUL + c) & 0xFFFFFFFFFFFFFFFF
        out.append(s >> 32)
    return out


def esim_direction_table():
    """The generator's 2048 directions as float64 (cos, sin) pairs with 21 significant bits: entry k is the double whose
    high word is the high word of cos/sin((2k+1)*pi/2048) rounded to nearest at bit 32 and whose low word is zero
    (tools/gen_dir_table.py builds the library's copy the same way; a test compares the two)."""
    t = (2 * np.arange(2048, dtype=np.float64) + 1) * (np.pi / 2048)
    cs = np.stack([np.cos(t), np.sin(t)], axis=1)
    bits = cs.view(np.uint64)
    return (((bits + np.uint64(0x80000000)) >> np.uint64(32)) << np.uint64(32)).view(np.float64)


def esim_noise_from_words(words, std, pixel_group, table=None):
    """Documented word -> noise mapping of the in-kernel generator, evaluated in float64 (the device evaluates the radius
    with the SFU's lg2 / sqrt approximations in float32, so this is a close, not a bit-exact, restatement): interval i
    uses words 2i and 2i+1; a word gives the Box-Muller pair of two neighbouring pixels (word 2i -> pixels 0,1, word
    2i+1 -> pixels 2,3): radius from its low 21 bits, u = (2^21 - low21) / 2^21, r = std*sqrt(-2 ln u); direction index
    = (top 8 bits << 3) | (pixel_group & 7); even pixel r*cos, odd pixel r*sin.  Returns [len(words)//2, 4]."""
    table = esim_direction_table() if table is None else table
    w = np.asarray(words, dtype=np.uint64)
    u = (2.0 ** 21 - (w & np.uint64(0x1FFFFF)).astype(np.float64)) / 2.0 ** 21
    r = np.float64(np.float32(std)) * np.sqrt(-2.0 * np.log(u))
    idx = (((w >> np.uint64(21)) & np.uint64(0x7F8)) | np.uint64(pixel_group & 7)).astype(np.int64)
    pairs = r[:, None] * table[idx]                       # [n_words, 2]
    return pairs.reshape(-1, 4)


# ---- voxel-space augmentation of the cached-voxel dataset -----------------------------------

def add_noise_to_voxel(voxel, noise_std=1.0, noise_fraction=0.1, integer_noise=False, rs=np.random):
    """data/esim_dataset.py:33-46.  Draw order on the legacy NumPy stream: integer noise = poisson(shape) then
    randint(0,2,shape); Gaussian = randn(shape); then rand(shape) for the mask when noise_fraction < 1.  Returns
    float64 ``voxel + noise`` (the caller's float32 array rounds it on assignment)."""
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * noise_std ** 2)) / 2
        y = rs.poisson(lam=lmb, size=voxel.shape)
        sign = 2 * rs.randint(0, 2, size=voxel.shape) - 1
        noise = y * sign
    else:
        noise = noise_std * rs.randn(*voxel.shape)
    if noise_fraction < 1.0:
        mask = rs.rand(*voxel.shape) >= noise_fraction
        noise = np.where(mask, 0, noise)
    return voxel + noise


def hot_pixel_draws(H, W, hot_pixel_std, max_hot_pixel_fraction, integer_noise, rs=np.random, pyrand=None):
    """The host draws of add_hot_pixels_to_voxels (data/esim_dataset.py:10-24): ``random.uniform`` for the fraction,
    randint x (columns), randint y (rows), then the values.  With integer noise the reference REUSES the name ``y`` for
    the Poisson draws (:19), so the rows the noise lands on are the Poisson values themselves — restated as is.
    Returns (rows, cols, values)."""
    import random as _random
    pyrand = _random if pyrand is None else pyrand
    frac = pyrand.uniform(0, max_hot_pixel_fraction)
    num = int(frac * H * W)
    x = rs.randint(0, W, num)
    y = rs.randint(0, H, num)
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * hot_pixel_std ** 2)) / 2
        y = rs.poisson(lam=lmb, size=num)
        sign = 2 * rs.randint(0, 2, size=num) - 1
        val = y * sign
    else:
        val = rs.randn(num)
        val *= hot_pixel_std
    return y, x, val


def add_hot_pixels_to_voxels(voxels, hot_pixel_std=1.0, max_hot_pixel_fraction=0.001, integer_noise=False, rs=np.random,
                             pyrand=None):
    """data/esim_dataset.py:7-30: one [H,W] noise map (np.add.at, float64) added in place to every [T,C] plane."""
    T, C, H, W = voxels.shape
    y, x, val = hot_pixel_draws(H, W, hot_pixel_std, max_hot_pixel_fraction, integer_noise, rs, pyrand)
    noise = np.zeros((H, W))
    np.add.at(noise, (y, x), val)
    voxels += noise[np.newaxis, np.newaxis, ...]
    return voxels


def cached_sequence_item(all_frame, all_flow, all_voxel, sequence_length, proba_pause_when_running, proba_pause_when_paused,
                         noise_std, noise_fraction, hot_pixel_std, max_hot_pixel_fraction, integer_noise, rs=np.random,
                         pyrand=None):
    """ESIMH5Dataset.__getitem__ after the crop / flip (data/esim_dataset.py:108-143): the pause sequence (one
    ``rand()`` per step, drawn BEFORE that step's voxel noise; a paused step repeats the previous frame and leaves flow
    and voxel zero), per-step add_noise_to_voxel, then add_hot_pixels_to_voxels on the whole sequence.  float32 in,
    float32 out (the reference's arrays are the h5 file's float32).  Returns (frame, flow, voxel, source_index) where
    source_index[t] is the cached sample used at step t, or -1 for a paused step."""
    frame, flow, voxel = np.zeros_like(all_frame), np.zeros_like(all_flow), np.zeros_like(all_voxel)
    src = np.full(sequence_length, -1, dtype=np.int64)
    paused, k = False, 0
    for t in range(sequence_length):
        u = rs.rand()
        paused = u < (proba_pause_when_paused if paused else proba_pause_when_running)
        if t > 0 and paused:
            frame[t] = frame[t - 1]
        else:
            frame[t], flow[t], voxel[t] = all_frame[k], all_flow[k], all_voxel[k]
            src[t] = k
            k += 1
        voxel[t] = add_noise_to_voxel(voxel[t], noise_std, noise_fraction, integer_noise, rs)
    voxel = add_hot_pixels_to_voxels(voxel, hot_pixel_std, max_hot_pixel_fraction, integer_noise, rs, pyrand)
    return frame, flow, voxel, src
