"""GPU parity of the event-stream scatter kernels (golden vectors of the reference + CPU oracle)."""
import os

import numpy as np
import pytest
import torch

import v2v_oracle as orc
from conftest import golden

pytestmark = pytest.mark.gpu

TOL = dict(rtol=1e-5, atol=1e-5)     # |a-b| <= 1e-5*max(1,|b|), BASELINE north_star


@pytest.mark.parametrize("name", golden("scatter").names("scat_mv_"))
def test_make_voxel_golden(cuda_device, name):
    import v2v_b200 as v2v
    c = golden("scatter").case(name)
    got = v2v.make_voxel([c["ts"], c["xs"], c["ys"], c["ps"]], int(c["bins"]), int(c["H"]), int(c["W"]), bool(c["interp"]))
    assert got.dtype == np.float64 and got.shape == c["ref"].shape
    if bool(c["interp"]):
        assert np.allclose(got, c["ref"], **TOL)
        assert np.max(np.abs(got - c["ref"])) < 1e-6
    else:
        assert np.array_equal(got, c["ref"])         # integer counts: exact


@pytest.mark.parametrize("name", golden("scatter").names("scat_tv_"))
def test_events_to_voxel_torch_golden(cuda_device, name):
    import v2v_b200 as v2v
    c = golden("scatter").case(name)
    hw = (int(c["H"]), int(c["W"]))
    args = [torch.from_numpy(c[k]) for k in ("xs", "ys", "ts", "ps")]
    got = v2v.events_to_voxel_torch(*args, int(c["bins"]), sensor_size=hw, temporal_bilinear=bool(c["bilinear"]))
    assert got.is_cuda and got.dtype == torch.float32
    assert np.allclose(got.cpu().numpy(), c["ref"], **TOL)
    p, n = v2v.events_to_neg_pos_voxel_torch(*args, int(c["bins"]), sensor_size=hw, temporal_bilinear=bool(c["bilinear"]))
    assert np.allclose(p.cpu().numpy(), c["ref_pos"], **TOL) and np.allclose(n.cpu().numpy(), c["ref_neg"], **TOL)
    if not bool(c["bilinear"]):
        assert np.array_equal(got.cpu().numpy(), c["ref"])


@pytest.mark.parametrize("name", ["scat_img_bil_pad", "scat_img_nearest_clip_nopad", "scat_img_nearest_default"])
def test_events_to_image_torch_golden(cuda_device, name):
    import v2v_b200 as v2v
    c = golden("scatter").case(name)
    got = v2v.events_to_image_torch(torch.from_numpy(c["xs"]), torch.from_numpy(c["ys"]), torch.from_numpy(c["ps"]),
                                    sensor_size=(int(c["H"]), int(c["W"])), clip_out_of_range=bool(c["clip"]),
                                    interpolation="bilinear" if int(c["bilinear"]) else None, padding=bool(c["padding"]))
    assert got.shape == c["ref"].shape
    assert np.allclose(got.cpu().numpy(), c["ref"], **TOL)


def test_events_to_image_np_and_count_map(cuda_device):
    import v2v_b200 as v2v
    c = golden("scatter").case("scat_img_np")
    got = v2v.events_to_image(c["xs"], c["ys"], c["ps"], sensor_size=(int(c["H"]), int(c["W"])))
    assert np.allclose(got, c["ref"], rtol=1e-12, atol=1e-12)
    cm = v2v.event_count_map(c["xs"], c["ys"], int(c["H"]), int(c["W"])).cpu().numpy()
    assert np.array_equal(cm, orc.event_count_map(c["xs"], c["ys"], int(c["H"]), int(c["W"])))


def synth_stream(ne, h, w, wn, seed, hot_frac=0.0):
    g = np.random.Generator(np.random.PCG64(seed))
    xs = g.integers(0, w, ne).astype(np.uint16)
    ys = g.integers(0, h, ne).astype(np.uint16)
    if hot_frac:
        sel = g.random(ne) < hot_frac
        xs[sel], ys[sel] = 3, 5
    ts = np.sort(g.random(ne) * 1.0 + 100.0)
    ps = (g.random(ne) < 0.5).astype(np.uint8)
    cuts = np.sort(g.integers(0, ne + 1, wn - 1))
    off = np.concatenate([[0], cuts, [ne]]).astype(np.int64)
    return ts, xs, ys, ps, off


@pytest.mark.parametrize("bins,interp", [(5, False), (5, True), (15, False), (15, True)])
def test_windows_vs_oracle(cuda_device, bins, interp):
    """HQF/MVSEC-shaped sensor, many ragged windows (some empty) in one launch; hot pixel included."""
    import v2v_b200 as v2v
    h, w, wn = 260, 346, 23
    ts, xs, ys, ps, off = synth_stream(200_000, h, w, wn, 17, hot_frac=0.01)
    off[5] = off[4]                                   # an empty window
    off = np.sort(off)
    got, dropped, unsorted = v2v.voxelize_windows(xs, ys, ts, ps, off, bins, h, w, mode="h5_interp" if interp else "h5_discrete",
                                                  return_dropped=True)
    assert int(dropped) == 0 and int(unsorted) == 0
    got = got.cpu().numpy()
    for k in range(wn):
        s = slice(off[k], off[k + 1])
        ref = orc.make_voxel(ts[s], xs[s], ys[s], ps[s], bins, h, w, interp)
        if interp:
            assert np.allclose(got[k], ref, **TOL)
        else:
            assert np.array_equal(got[k], ref.astype(np.float32))


def test_scatter_full_size_properties(cuda_device):
    """BASELINE config 4 size (10 M events, 260x346, 400 windows): conservation instead of a full oracle run."""
    import v2v_b200 as v2v
    h, w, wn, ne = 260, 346, 400, 10_000_000
    ts, xs, ys, ps, off = synth_stream(ne, h, w, wn, 23, hot_frac=0.01)
    pol = 2 * ps.astype(np.int64) - 1
    csum = np.concatenate([[0], np.cumsum(pol)])
    per_win = csum[off[1:]] - csum[off[:-1]]
    vd = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_discrete")
    assert np.array_equal(vd.sum(dim=(1, 2, 3)).cpu().numpy().astype(np.int64), per_win)
    vi = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_interp", out_dtype=torch.float64)
    assert np.allclose(vi.sum(dim=(1, 2, 3)).cpu().numpy(), per_win, atol=1e-2)
    # bins collapse: summing the interpolated voxel over bins equals the discrete one summed over bins (up to the guard)
    assert torch.allclose(vi.sum(dim=1), vd.sum(dim=1).to(torch.float64), atol=1e-3)
    # spot check a few windows against the oracle
    for k in (0, 137, 399):
        s = slice(off[k], off[k + 1])
        assert np.array_equal(vd[k].cpu().numpy(), orc.make_voxel(ts[s], xs[s], ys[s], ps[s], 5, h, w, False).astype(np.float32))
        assert np.allclose(vi[k].cpu().numpy(), orc.make_voxel(ts[s], xs[s], ys[s], ps[s], 5, h, w, True), **TOL)


def test_out_of_sensor_events_are_dropped_and_counted(cuda_device):
    import v2v_b200 as v2v
    ts = np.linspace(0, 0.01, 6)
    xs = np.array([0, 1, 50, 2, 3, 4], dtype=np.int16)
    ys = np.array([0, 1, 1, -1, 3, 40], dtype=np.int16)
    ps = np.ones(6, dtype=np.uint8)
    v, dropped, _ = v2v.voxelize_windows(xs, ys, ts, ps, [0, 6], 5, 8, 8, return_dropped=True)
    assert int(dropped) == 3 and float(v.sum()) == 3.0


def test_mixin_matches_reference_signature(cuda_device):
    import v2v_b200 as v2v
    c = golden("scatter").case("scat_mv_disc5")

    class DS(v2v.MakeVoxelMixin):
        num_bins, H, W, interpolate_bins = int(c["bins"]), int(c["H"]), int(c["W"]), False
    assert np.array_equal(DS().make_voxel([c["ts"], c["xs"], c["ys"], c["ps"]]), c["ref"])


def test_large_bins_take_the_32bit_fallback(cuda_device):
    """More than 32767 events in one bin of one window (and half of them on one pixel): the packed 16-bit
    counters are not usable, the kernel must fall back to two 32-bit passes and stay exact."""
    import v2v_b200 as v2v
    h, w, ne = 64, 96, 400_000
    g = np.random.Generator(np.random.PCG64(9))
    xs = g.integers(0, w, ne).astype(np.uint16)
    ys = g.integers(0, h, ne).astype(np.uint16)
    hot = g.random(ne) < 0.5
    xs[hot], ys[hot] = 7, 63
    ts = np.sort(g.random(ne)) * 0.2 + 1.0
    ps = (g.random(ne) < 0.5).astype(np.uint8)
    ps[hot] = 1                                           # hot pixel: > 32767 positive events per bin
    for bins in (5, 3):
        got = v2v.make_voxel([ts, xs, ys, ps], bins, h, w, False)
        ref = orc.make_voxel(ts, xs, ys, ps, bins, h, w, False)
        assert np.abs(ref).max() > 32767
        assert np.array_equal(got, ref)
    gi = v2v.make_voxel([ts, xs, ys, ps], 5, h, w, True)
    ri = orc.make_voxel(ts, xs, ys, ps, 5, h, w, True)
    assert np.allclose(gi, ri, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["h5_discrete", "h5_interp", "torch_discrete", "torch_bilinear"])
def test_split_bins_across_ctas(cuda_device, mode):
    """Large single window: every bin's events are sliced over several CTAs and combined with global atomics.
    Forced here through the tuning knob, compared with the unsplit result and the oracle."""
    import v2v_b200 as v2v
    h, w, ne = 90, 120, 300_000
    g = np.random.Generator(np.random.PCG64(12))
    xs = g.integers(0, w, ne).astype(np.int16)
    ys = g.integers(0, h, ne).astype(np.int16)
    ts = np.sort(g.random(ne)) * 0.5
    if mode.startswith("h5"):
        ps = (g.random(ne) < 0.5).astype(np.uint8)
        tsx = ts + 3.0
    else:
        ps = (g.random(ne) < 0.5).astype(np.float32) * 2 - 1
        tsx = ts.astype(np.float32)
    base = v2v.voxelize_windows(xs, ys, tsx, ps, [0, ne], 5, h, w, mode=mode).cpu().numpy()
    os.environ["V2V_SCATTER_SPLITS"] = "7"
    try:
        split, dropped, _ = v2v.voxelize_windows(xs, ys, tsx, ps, [0, ne], 5, h, w, mode=mode, return_dropped=True)
    finally:
        del os.environ["V2V_SCATTER_SPLITS"]
    assert int(dropped) == 0
    split = split.cpu().numpy()
    if mode.endswith("discrete"):
        assert np.array_equal(split, base)
    else:
        assert np.allclose(split, base, **TOL)
    if mode == "h5_discrete":
        assert np.array_equal(split[0], orc.make_voxel(tsx, xs, ys, ps, 5, h, w, False).astype(np.float32))
    if mode == "h5_interp":
        assert np.allclose(split[0], orc.make_voxel(tsx, xs, ys, ps, 5, h, w, True), **TOL)


def test_full_size_stream_vs_c_oracle(cuda_device):
    """BASELINE config 4 (10 M events, 260x346, 400 windows): every window against the C oracle."""
    import v2v_b200 as v2v
    import v2v_oracle_c as orcc
    h, w, wn, ne = 260, 346, 400, 10_000_000
    ts, xs, ys, ps, off = synth_stream(ne, h, w, wn, 29, hot_frac=0.01)
    vd = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_discrete").cpu().numpy()
    vi = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_interp", out_dtype=torch.float64).cpu().numpy()
    worst = 0.0
    for k in range(wn):
        s = slice(off[k], off[k + 1])
        assert np.array_equal(vd[k], orcc.make_voxel(ts[s], xs[s], ys[s], ps[s], 5, h, w, False).astype(np.float32)), k
        ri = orcc.make_voxel(ts[s], xs[s], ys[s], ps[s], 5, h, w, True)
        worst = max(worst, float(np.max(np.abs(vi[k] - ri) / np.maximum(1.0, np.abs(ri)))))
    assert worst <= 1e-5, worst
    assert worst < 1e-6          # in practice the fixed-point accumulation is ~1e-8


def test_fps_windows_and_raw_event_packing(cuda_device):
    """FPS_H5Dataset window segmentation (data/testh5.py:468-474) and the NER-Net raw event tensor
    (:329-339), both exact against NumPy."""
    import v2v_b200 as v2v
    g = np.random.Generator(np.random.PCG64(77))
    ne, h, w = 200_000, 180, 240
    ts = np.sort(g.random(ne)) * 2.37 + 11.0
    ts[1000:1010] = ts[1000]                                  # ties
    xs = g.integers(0, w, ne).astype(np.uint16)
    ys = g.integers(0, h, ne).astype(np.uint16)
    ps = (g.random(ne) < 0.5).astype(np.uint8)
    fps = 100
    total = int((ts[-1] - ts[0]) * fps)
    borders = np.linspace(ts[0], ts[-1], total + 1)
    ref_idx = np.searchsorted(ts, borders)
    idx, b2 = v2v.fps_window_offsets(ts, fps)
    assert np.array_equal(b2, borders) and np.array_equal(idx.cpu().numpy(), ref_idx)
    # the windows feed the scatter directly: whole sequence in one launch == per-window reference voxels
    vox = v2v.voxelize_windows(xs, ys, ts, ps, idx, 5, h, w, mode="h5_discrete").cpu().numpy()
    for k in (0, total // 2, total - 1):
        s = slice(ref_idx[k], ref_idx[k + 1])
        assert np.array_equal(vox[k], orc.make_voxel(ts[s], xs[s], ys[s], ps[s], 5, h, w, False).astype(np.float32))
    s = slice(ref_idx[3], ref_idx[4])
    ref = np.stack([xs[s].astype(np.float64), ys[s].astype(np.float64), ts[s].astype(np.float64),
                    ps[s].astype(np.float64) * 2 - 1, np.zeros(s.stop - s.start)], axis=1)
    got = v2v.pack_events_n5(xs[s], ys[s], ts[s], ps[s])
    assert got.dtype == torch.float64 and np.array_equal(got.cpu().numpy(), ref)
    assert v2v.pack_events_n5(xs[:0], ys[:0], ts[:0], ps[:0]).shape == (1, 5)


def test_validate_flag_rejects_unsorted_windows(cuda_device):
    import v2v_b200 as v2v
    ts = np.array([0.0, 0.1, 0.2, 0.05, 0.06, 0.07])          # decrease at index 3 = window boundary: allowed
    xs = np.arange(6, dtype=np.int16)
    ys = np.zeros(6, dtype=np.int16)
    ps = np.ones(6, dtype=np.uint8)
    v = v2v.voxelize_windows(xs, ys, ts, ps, [0, 3, 6], 2, 4, 8, validate=True)
    assert float(v.sum()) == 6.0
    with pytest.raises(ValueError):
        v2v.voxelize_windows(xs, ys, ts, ps, [0, 6], 2, 4, 8, validate=True)


def test_numpy_events_to_voxel_golden(cuda_device):
    import v2v_b200 as v2v
    c = golden("scatter").case("scat_voxel_np")
    got = v2v.events_to_voxel(c["xs"], c["ys"], c["ts"][:, None], c["ps"][:, None], int(c["bins"]),
                              sensor_size=(int(c["H"]), int(c["W"])))
    assert got.dtype == np.float64 and np.allclose(got, c["ref"], rtol=1e-9, atol=1e-9)
    with pytest.raises(NotImplementedError):
        v2v.events_to_voxel(c["xs"], c["ys"], c["ts"], c["ps"], 5, temporal_bilinear=False)


def test_interp_paths_agree_and_take_any_event_order(cuda_device):
    """The interpolated mode has two kernels: the one-visit path (counting sort by strip; default whenever the workspace
    is there) and the contiguous-range path.  Both equal the oracle on sorted windows; on a SHUFFLED window the one-visit
    path still equals the reference (np.add.at does not care about order, data/testh5.py:74-80, as long as the first and
    the last event stay in place: they define the time axis, :68,76), because every event's bin comes from its own
    timestamp."""
    import v2v_b200 as v2v
    from v2v_b200 import _lib
    h, w, wn = 260, 346, 9
    ts, xs, ys, ps, off = synth_stream(120_000, h, w, wn, 5, hot_frac=0.02)
    a = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_interp", out_dtype=torch.float64)
    b = v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_interp", out_dtype=torch.float64,
                             kernel_flags=_lib.SCATTER_FLAG_RANGES)
    assert torch.allclose(a, b, rtol=0, atol=2e-6)          # (items of <= 255 records accumulate in 2^-23 units)
    for k in range(wn):
        s = slice(off[k], off[k + 1])
        assert np.allclose(a[k].cpu().numpy(), orc.make_voxel(ts[s], xs[s], ys[s], ps[s], 5, h, w, True), **TOL)
    # shuffle the interior of every window
    g = np.random.Generator(np.random.PCG64(1))
    perm = np.arange(len(ts))
    for k in range(wn):
        if off[k + 1] - off[k] > 2:
            seg = perm[off[k] + 1: off[k + 1] - 1]
            g.shuffle(seg)
    tsu, xsu, ysu, psu = ts[perm], xs[perm], ys[perm], ps[perm]
    u = v2v.voxelize_windows(xsu, ysu, tsu, psu, off, 5, h, w, mode="h5_interp", out_dtype=torch.float64)
    assert torch.equal(u, a)                                # integer accumulation: independent of the event order
    for k in (0, 4):
        s = slice(off[k], off[k + 1])
        assert np.allclose(u[k].cpu().numpy(), orc.make_voxel(tsu[s], xsu[s], ysu[s], psu[s], 5, h, w, True), **TOL)


def test_interp_hot_pixel_beyond_the_two_word_range(cuda_device):
    """300 k same-sign events on ONE pixel of one window, 3 bins: the middle bin's cell sums ~150 k unit weights, more than
    the two 32-bit fixed-point words of a cell hold (65535).  Both interpolated kernels must switch that item to the
    64-bit accumulator and stay within n * 2^-31 of the reference (data/testh5.py:74-80)."""
    import v2v_b200 as v2v
    from v2v_b200 import _lib
    h, w, ne = 40, 64, 320_000
    g = np.random.Generator(np.random.PCG64(21))
    xs = g.integers(0, w, ne).astype(np.uint16)
    ys = g.integers(0, h, ne).astype(np.uint16)
    hot = g.random(ne) < 0.94
    xs[hot], ys[hot] = 11, 17
    ts = np.sort(g.random(ne)) * 0.5 + 3.0
    ps = (g.random(ne) < 0.5).astype(np.uint8)
    ps[hot] = 0                                           # negative hot pixel: the sign path of the wide form
    off = np.array([0, ne], dtype=np.int64)
    ref = orc.make_voxel(ts, xs, ys, ps, 3, h, w, True)
    assert np.abs(ref).max() > 100_000
    for flags in (0, _lib.SCATTER_FLAG_RANGES):
        got = v2v.voxelize_windows(xs, ys, ts, ps, off, 3, h, w, mode="h5_interp", out_dtype=torch.float64,
                                   kernel_flags=flags)[0].cpu().numpy()
        assert np.abs(got - ref).max() < 2e-4, (flags, np.abs(got - ref).max())
    got32 = v2v.make_voxel([ts, xs, ys, ps], 3, h, w, True)
    assert np.allclose(got32, ref, rtol=1e-6, atol=1e-3)


def test_unsorted_windows_raise_in_contiguous_range_modes(cuda_device):
    """The discrete and torch modes assign bins as contiguous ranges: unsorted timestamps must not silently mis-bin.
    The device-side check (default on) raises; a decrease exactly at a window boundary is legitimate."""
    import v2v_b200 as v2v
    h, w = 32, 40
    ts, xs, ys, ps, off = synth_stream(5000, h, w, 4, 9)
    v2v.voxelize_windows(xs, ys, ts, ps, off, 5, h, w, mode="h5_discrete")             # sorted: fine
    ts2 = ts.copy()
    ts2[off[1]:off[2]] -= 10.0                                                        # window 1 starts earlier than window 0 ends
    v2v.voxelize_windows(xs, ys, ts2, ps, off, 5, h, w, mode="h5_discrete")
    bad = ts.copy()
    i = off[2] + 7
    bad[i], bad[i + 1] = bad[i + 1], bad[i]
    assert bad[i] != bad[i + 1]
    for mode in ("h5_discrete", "torch_discrete", "torch_bilinear"):
        t = bad if mode.startswith("h5") else bad.astype(np.float32)
        p = ps if mode.startswith("h5") else ps.astype(np.float32) * 2 - 1
        with pytest.raises(ValueError):
            v2v.voxelize_windows(xs, ys, t, p, off, 5, h, w, mode=mode)
        _, _, unsorted = v2v.voxelize_windows(xs, ys, t, p, off, 5, h, w, mode=mode, return_dropped=True)
        assert int(unsorted) == 1
    v2v.voxelize_windows(xs, ys, bad, ps, off, 5, h, w, mode="h5_interp")              # any order


def test_neg_pos_voxel_single_launch(cuda_device):
    """events_to_neg_pos_voxel_torch is one launch (V2V_POL_SPLIT) and equals the two one-polarity launches."""
    import v2v_b200 as v2v
    g = np.random.Generator(np.random.PCG64(12))
    ne, h, w = 30_000, 60, 80
    xs, ys = g.integers(0, w, ne).astype(np.float32), g.integers(0, h, ne).astype(np.float32)
    ts = np.sort(g.random(ne)).astype(np.float32)
    ps = ((g.random(ne) < 0.5).astype(np.float32) * 2 - 1)
    for bil in (True, False):
        n0 = v2v.launch_count()
        pos, neg = v2v.events_to_neg_pos_voxel_torch(xs, ys, ts, ps, 5, sensor_size=(h, w), temporal_bilinear=bil)
        launches = v2v.launch_count() - n0
        mode = "torch_bilinear" if bil else "torch_discrete"
        p1 = v2v.voxelize_windows(xs, ys, ts, ps, [0, ne], 5, h, w, mode=mode, polarity="pos")[0]
        n1 = v2v.voxelize_windows(xs, ys, ts, ps, [0, ne], 5, h, w, mode=mode, polarity="neg")[0]
        assert torch.allclose(pos, p1, **TOL) and torch.allclose(neg, n1, **TOL)
        assert launches <= 3                         # sortedness check + bin boundaries + ONE scatter launch
        rp, rn = orc.events_to_neg_pos_voxel_f32(xs, ys, ts, ps, 5, (h, w), bil)
        assert np.allclose(pos.cpu().numpy(), rp, **TOL) and np.allclose(neg.cpu().numpy(), rn, **TOL)
