"""GPU twins of test_properties_cpu.py: the kernels against the oracle on RANDOM shapes, thresholds, bin layouts and event
streams drawn by hypothesis (ragged planes, single rows / columns, 2-frame clips, empty windows ...), through the C ABI.
Counts and discrete voxels bit-exact; interpolated voxels within the fixed-point bound."""
import os

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import v2v_oracle as orc

pytestmark = pytest.mark.gpu
# 60 fixed examples per test by default (the same on every run); V2V_PROP_EXAMPLES=N runs N fresh random ones (bug hunts)
_N = int(os.environ.get("V2V_PROP_EXAMPLES", "0"))
SET = dict(max_examples=_N or 60, deadline=None, derandomize=not _N, suppress_health_check=[HealthCheck.function_scoped_fixture])


def _video(rs, n, h, w):
    base = rs.randint(0, 256, (h, w)).astype(np.int64)
    steps = rs.randint(-40, 41, (n, h, w))
    steps[0] = 0
    return np.clip(base[None] + np.cumsum(steps, 0), 0, 255).astype(np.uint8)


@given(seed=st.integers(0, 2 ** 31 - 1), t=st.integers(1, 3), bins=st.sampled_from([1, 2, 5]), fpb=st.sampled_from([1, 2]),
       h=st.integers(1, 12), w=st.integers(1, 16), pos=st.floats(0.05, 1.5), gap=st.floats(1.0, 1.5), std=st.floats(0.0, 0.3),
       external=st.booleans(), flags=st.sampled_from([0, 1, 2, 4]))
@settings(**SET)
def test_esim_random_shapes_equal_oracle(cuda_device, seed, t, bins, fpb, h, w, pos, gap, std, external, flags):
    """Explicit random fields, any plane shape and bin layout, every kernel choice (library's, generic, throughput kernel on
    small launches, one pixel per thread): the crossing counts equal the oracle's bit for bit, the statistics equal their
    sums, and with external noise the voxels are the float32 rounding of the oracle's float64 values."""
    import v2v_b200 as v2v
    rs = np.random.RandomState(seed)
    n = t * bins * fpb + 1
    neg = pos * gap
    video = _video(rs, n, h, w)
    u0, hot, g = orc.esim_draw_randomness(n, h, w, 0.2, 0.5, rs)
    per_interval = orc.esim_video_to_voxel(video, pos, neg, std, u0, hot, g, external)
    ref = orc.bin_accumulate(per_interval, bins, fpb)
    o = v2v.frames_to_voxel(torch.from_numpy(video).to(cuda_device), pos, neg, num_bins=bins, frames_per_bin=fpb, noise="explicit",
                            base_noise_std=std, put_noise_external=external, u0=u0[None], hot_noise=hot[None], base_gauss=g[None],
                            with_stats=not external, kernel_flags=flags)
    got = o.voxel[0].cpu().numpy()
    assert got.shape == ref.shape
    if external:
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    else:
        assert np.array_equal(got.astype(np.float64), ref)
        s = o.stats[0].cpu().numpy()
        # (event totals are counted per interval, before a bin nets opposite signs)
        assert s[0] == int(np.maximum(per_interval, 0).sum()) and s[1] == int(np.maximum(-per_interval, 0).sum())


@given(seed=st.integers(0, 2 ** 31 - 1), ne=st.integers(0, 3000), wn=st.integers(1, 6), h=st.integers(1, 40), w=st.integers(1, 50),
       bins=st.sampled_from([1, 2, 5, 15]), interp=st.booleans(), f32=st.booleans(), ranges=st.booleans())
@settings(**SET)
def test_scatter_random_streams_equal_oracle(cuda_device, seed, ne, wn, h, w, bins, interp, f32, ranges):
    """Random sorted streams cut into random windows (empty ones included): every window equals TestH5Dataset.make_voxel's
    restatement (data/testh5.py:60-90) — discrete exact, interpolated within the fixed-point bound, on both interpolated paths."""
    import v2v_b200 as v2v
    from v2v_b200 import _lib
    rs = np.random.RandomState(seed)
    ts = np.sort(rs.rand(ne) * 0.2 + 7.0).astype(np.float32 if f32 else np.float64)
    xs = rs.randint(0, w, ne).astype(np.uint16)
    ys = rs.randint(0, h, ne).astype(np.uint16)
    ps = rs.randint(0, 2, ne).astype(np.uint8)
    off = np.sort(np.concatenate([[0, ne], rs.randint(0, ne + 1, wn - 1)])).astype(np.int64)
    got = v2v.voxelize_windows(xs, ys, ts, ps, off, bins, h, w, mode="h5_interp" if interp else "h5_discrete", out_dtype=torch.float64,
                               kernel_flags=_lib.SCATTER_FLAG_RANGES if ranges else 0).cpu().numpy()
    assert got.shape == (wn, bins, h, w)
    for k in range(wn):
        s = slice(off[k], off[k + 1])
        if off[k + 1] == off[k]:
            assert not got[k].any()
            continue
        ref = orc.make_voxel(ts[s], xs[s], ys[s], ps[s], bins, h, w, interp)
        if interp:
            # the documented bound: weights are rounded to 2^-23 once (|error| <= 2^-24 per event on a cell) in work items of
            # at most 255 events, to 2^-30 in larger ones
            nwin = int(off[k + 1] - off[k])
            assert np.allclose(got[k], ref, rtol=0, atol=max(2e-6 * max(1.0, np.abs(ref).max()), min(nwin, 255) * 2.0 ** -24 + nwin * 2.0 ** -31))
        else:
            assert np.array_equal(got[k], ref)


@given(seed=st.integers(0, 2 ** 31 - 1), t=st.integers(1, 4), bins=st.sampled_from([1, 5]), h=st.integers(1, 24), w4=st.integers(1, 16),
       ragged=st.booleans(), pos=st.floats(0.05, 1.0), gap=st.floats(1.0, 1.5), std=st.floats(0.0, 0.1), frac=st.sampled_from([0.0, 0.001, 0.05]),
       hstd=st.floats(0.0, 10.0), base=st.integers(0, 1 << 20), flags=st.sampled_from([0, 1, 2, 4, 16]))
@settings(**SET)
def test_esim_production_mode_random_shapes(cuda_device, seed, t, bins, h, w4, ragged, pos, gap, std, frac, hstd, base, flags):
    """The production noise mode (in-kernel generator) on random shapes and parameters: whatever kernel is chosen (library's
    choice, generic, throughput kernel on small launches, one pixel per thread, bulk-copy ring) the run equals the CPU oracle
    fed with the fields the generator drew for that (seed, global clip index) — dumped through the audit hook — bit for bit,
    including the residual potential and the event statistics."""
    import v2v_b200 as v2v
    rs = np.random.RandomState(seed)
    w = 4 * w4 + (int(rs.randint(1, 4)) if ragged else 0)
    n = t * bins + 1
    neg = pos * gap
    video = _video(rs, n, h, w)
    o = v2v.frames_to_voxel(torch.from_numpy(video).to(cuda_device), pos, neg, num_bins=bins, noise="philox", base_noise_std=std,
                            hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=seed, clip_index_base=base, with_stats=True,
                            return_potential=True, kernel_flags=flags)
    u0, hot, bn = v2v.philox_fields(n, h, w, base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=seed, clip_index_base=base)
    u0, hot, bn = u0[0].cpu().numpy(), hot[0].cpu().numpy(), bn[0].cpu().numpy()
    assert np.isfinite(bn).all() and np.isfinite(hot).all() and 0.0 <= u0.min() and u0.max() < 1.0
    per_interval, pot = orc.esim_video_to_voxel(video, pos, neg, 1.0, u0, hot, bn, False, return_state=True)
    assert np.array_equal(o.voxel[0].cpu().numpy().astype(np.float64), orc.bin_accumulate(per_interval, bins, 1))
    assert np.array_equal(o.potential[0].cpu().numpy(), pot)
    s = o.stats[0].cpu().numpy()
    assert s[0] == int(np.maximum(per_interval, 0).sum()) and s[1] == int(np.maximum(-per_interval, 0).sum())


@given(seed=st.integers(0, 2 ** 31 - 1), ne=st.integers(2, 4000), h=st.integers(1, 30), w=st.integers(1, 40), bins=st.sampled_from([1, 3, 5, 9]),
       bilinear=st.booleans())
@settings(**SET)
def test_torch_voxel_random_streams_equal_oracle(cuda_device, seed, ne, h, w, bins, bilinear):
    """events_to_voxel_torch / events_to_neg_pos_voxel_torch (utils/event_utils.py:466-541) on random streams in the dtypes its
    callers pass (float32 coordinates, timestamps relative to the first event, polarities +-1): discrete exact, bilinear
    within float32 summation order; the one-launch neg / pos split equals the two reference calls."""
    import v2v_b200 as v2v
    rs = np.random.RandomState(seed)
    ts = np.sort(rs.rand(ne) * 0.3).astype(np.float32)
    ts -= ts[0]
    if ts[-1] == 0:
        ts[-1] = np.float32(0.01)
    xs = rs.randint(0, w, ne).astype(np.float32)
    ys = rs.randint(0, h, ne).astype(np.float32)
    ps = (2.0 * rs.randint(0, 2, ne) - 1.0).astype(np.float32)
    ref = orc.events_to_voxel_f32(xs, ys, ts, ps, bins, (h, w), bilinear)
    rp, rn = orc.events_to_neg_pos_voxel_f32(xs, ys, ts, ps, bins, (h, w), bilinear)
    args = [torch.from_numpy(a) for a in (xs, ys, ts, ps)]
    got = v2v.events_to_voxel_torch(*args, bins, sensor_size=(h, w), temporal_bilinear=bilinear).cpu().numpy()
    gp, gn = v2v.events_to_neg_pos_voxel_torch(*args, bins, sensor_size=(h, w), temporal_bilinear=bilinear)
    tol = dict(rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(ref).max())))
    if bilinear:
        assert np.allclose(got, ref, **tol) and np.allclose(gp.cpu().numpy(), rp, **tol) and np.allclose(gn.cpu().numpy(), rn, **tol)
    else:
        assert np.array_equal(got, ref) and np.array_equal(gp.cpu().numpy(), rp) and np.array_equal(gn.cpu().numpy(), rn)


@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(2, 9), h=st.integers(1, 14), w=st.integers(1, 18), cutoff=st.sampled_from([0.0, 15.0, 200.0]),
       leak=st.sampled_from([0.0, 0.1, 2.0]), shot=st.sampled_from([0.0, 5.0, 40.0]), jitter=st.sampled_from([0.0, 0.2]),
       model=st.sampled_from(["pn_related", "spatial_independent"]), u8=st.booleans(), flags=st.sampled_from([0, 1, 2]))
@settings(**SET)
def test_v2e_random_presets_equal_oracle(cuda_device, seed, n, h, w, cutoff, leak, shot, jitter, model, u8, flags):
    """v2e-style core (data/v2v_core_v2e.py:401-581) on random shapes and feature combinations (low-pass, leak + jitter, shot
    noise, both time-invariant threshold models, uint8 or float video: the wrapping intensity rescale of :190): the kernels fed
    with the fields the oracle drew equal it bit for bit, whichever kernel runs."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    rs = np.random.RandomState(seed)
    video = _video(rs, n, h, w)
    video = np.clip((video.astype(np.float64) - 127.5) * 2.1 + 127.5, 0, 255).astype(np.uint8)        # HDR degrade: many values >= 236
    p = dict(threshold_model=model, thres_mean_mean=0.2, thres_mean_std=0.04, thres_diff_mean=0.02, thres_diff_std=0.03, cutoff_hz=cutoff,
             leak_rate_hz=leak, shot_noise_rate_hz=shot, leak_jitter_fraction=jitter, noise_rate_cov_decades=0.1)
    rec = {}
    ref = orc.v2e_video_to_voxel(video if u8 else video.astype(np.float64), 24, p, np.random.RandomState(seed + 1), record=rec)
    out = frames_to_voxel_v2e(torch.from_numpy(video).to(cuda_device), rec["pos_thres"][None], rec["neg_thres"][None], fps=24, cutoff_hz=cutoff,
                              leak_rate_hz=leak, shot_noise_rate_hz=shot, leak_jitter_fraction=jitter, noise_rate=rec["noise_rate"][None],
                              pos_thres_nominal=0.2 + 0.01, neg_thres_nominal=0.2 - 0.01, noise="explicit",
                              leak_randn=np.stack(rec["leak_randn"])[None] if rec["leak_randn"] else None,
                              pos_shot=np.stack(rec["pos_shot"]).astype(np.int32)[None] if rec["pos_shot"] else None,
                              neg_shot=np.stack(rec["neg_shot"]).astype(np.int32)[None] if rec["neg_shot"] else None,
                              u8_intensity=u8, kernel_flags=flags)
    assert np.array_equal(out["voxel"][0, :, 0].cpu().numpy().astype(np.float64), ref)
