"""CPU: the oracle restatement must reproduce the reference's golden outputs bit for bit."""
import numpy as np
import pytest

import v2v_oracle as orc
from conftest import golden


def same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", golden("esim").names("esim_"))
def test_esim_core(name):
    c = golden("esim").case(name)
    out = orc.esim_video_to_voxel(c["video"], float(c["pos"]), float(c["neg"]), float(c["base_noise_std"]),
                                  c["u0"], c["hot"], c["g"], bool(c["external"]), lut=c["lut"])
    assert same(out, c["ref"])


@pytest.mark.parametrize("name", golden("esim").names("esim_"))
def test_esim_draw_order(name):
    """Re-drawing from the same seed reproduces the stored random fields (draw order of the reference)."""
    c = golden("esim").case(name)
    n, h, w = c["video"].shape
    np.random.seed(int(c["seed"]))
    u0, hot, g = orc.esim_draw_randomness(n, h, w, float(c["hot_pixel_fraction"]), float(c["hot_pixel_std"]))
    assert same(u0, c["u0"]) and same(hot, c["hot"]) and same(g, c["g"])


@pytest.mark.parametrize("name", golden("esim").names("esimds_"))
def test_esim_dataset_level(name):
    c = golden("esim").case(name)
    iv = orc.esim_video_to_voxel(c["video"], float(c["p_pos_thres"]), float(c["p_neg_thres"]),
                                 float(c["p_base_noise_std"]), c["u0"], c["hot"], c["g"], bool(c["external"]),
                                 lut=c["lut"])
    out = orc.bin_accumulate(iv, int(c["bins"]), int(c["fpb"]))
    assert same(out, c["ref"])


def test_lut_matches_stored():
    c = golden("esim").case("esim_walk_clean")
    lut = orc.esim_log_lut()
    # identical on the host that generated the fixtures; 1-ulp differences are possible on other SIMD targets
    assert np.max(np.abs(lut - c["lut"]) / np.abs(c["lut"])) < 1e-15
    assert lut[0] == pytest.approx(-6.907755278982137) and lut[255] == pytest.approx(0.0009995003330834232)


def test_bin_accumulate_rejects_ragged():
    with pytest.raises(AssertionError):
        orc.bin_accumulate(np.zeros((7, 2, 2)), 5, 1)


@pytest.mark.parametrize("name", ["frames_add0", "frames_add1"])
def test_pack_frames(name):
    c = golden("esim").case(name)
    out = orc.pack_frames(c["imgs"], int(c["fpi"]), int(c["cnt"]), bool(c["add"]))
    assert same(out, c["ref"]) and out.dtype == np.float32


def v2e_params(c):
    p = {k[2:]: float(v) for k, v in c.items() if k.startswith("p_")}
    p["threshold_model"] = str(c["threshold_model"])
    return p


@pytest.mark.parametrize("name", golden("v2e").cases)
def test_v2e(name):
    c = golden("v2e").case(name)
    p = v2e_params(c)
    np.random.seed(int(c["seed"]))
    video = c["video"] if int(c.get("u8_input", 0)) else c["video"].astype(np.float64)   # uint8 input: wrapping inten01
    out = orc.v2e_video_to_voxel(video, int(c["fps"]), p, np.random, lut=c["lut"])
    assert same(out, c["ref"])
    if p["threshold_model"] == "spatial_temporal_independent":      # (the replay helper covers the time-invariant models)
        return
    fields = {k: c[k] for k in ("thr_a", "thr_b", "noise_randn")}
    for k in ("leak_randn", "pos_shot", "neg_shot"):
        fields[k] = list(c[k]) if k in c else []
    assert same(orc.v2e_replay(video, int(c["fps"]), p, fields, lut=c["lut"]), c["ref"])


@pytest.mark.parametrize("name", golden("scatter").names("scat_mv_"))
def test_make_voxel(name):
    c = golden("scatter").case(name)
    out = orc.make_voxel(c["ts"], c["xs"], c["ys"], c["ps"], int(c["bins"]), int(c["H"]), int(c["W"]),
                         bool(c["interp"]))
    assert same(out, c["ref"]) and out.dtype == np.float64


@pytest.mark.parametrize("name", golden("scatter").names("scat_tv_"))
def test_events_to_voxel_f32(name):
    c = golden("scatter").case(name)
    hw = (int(c["H"]), int(c["W"]))
    out = orc.events_to_voxel_f32(c["xs"], c["ys"], c["ts"], c["ps"], int(c["bins"]), hw, bool(c["bilinear"]))
    assert same(out, c["ref"]) and out.dtype == np.float32
    p, n = orc.events_to_neg_pos_voxel_f32(c["xs"], c["ys"], c["ts"], c["ps"], int(c["bins"]), hw, bool(c["bilinear"]))
    assert same(p, c["ref_pos"]) and same(n, c["ref_neg"])


@pytest.mark.parametrize("name", ["scat_img_bil_pad", "scat_img_nearest_clip_nopad", "scat_img_nearest_default"])
def test_events_to_image_f32(name):
    c = golden("scatter").case(name)
    out = orc.events_to_image_f32(c["xs"], c["ys"], c["ps"], (int(c["H"]), int(c["W"])), bool(c["clip"]),
                                  "bilinear" if int(c["bilinear"]) else None, bool(c["padding"]))
    assert same(out, c["ref"])


def test_events_to_image_np():
    c = golden("scatter").case("scat_img_np")
    assert same(orc.events_to_image_np(c["xs"], c["ys"], c["ps"], (int(c["H"]), int(c["W"]))), c["ref"])


def test_scatter_conservation():
    """Property: discrete voxels sum to the signed event count; interpolated ones to it up to the 1e-4 µs guard."""
    g = np.random.Generator(np.random.PCG64(3))
    n, h, w = 5000, 20, 30
    ts = np.sort(g.random(n)) * 0.03 + 4.0
    xs = g.integers(0, w, n).astype(np.uint16)
    ys = g.integers(0, h, n).astype(np.uint16)
    ps = (g.random(n) < 0.4).astype(np.uint8)
    v = orc.make_voxel(ts, xs, ys, ps, 5, h, w, False)
    assert v.sum() == (2 * ps.astype(np.int64) - 1).sum()
    vi = orc.make_voxel(ts, xs, ys, ps, 5, h, w, True)
    assert abs(vi.sum() - v.sum()) < 1e-3


def test_esim_properties():
    """Conservation of the integrator and single-polarity per pixel-interval."""
    from conftest import synth_video
    vid = synth_video("iid", 9, 12, 10, 5)
    lut = orc.esim_log_lut()
    pos, neg = 0.17, 0.23
    u0 = np.random.Generator(np.random.PCG64(1)).random((12, 10))
    z = np.zeros((12, 10))
    out, pot = orc.esim_video_to_voxel(vid, pos, neg, 0.0, u0, z, np.zeros((8, 12, 10)), False, lut, return_state=True)
    pe, ne = np.maximum(out, 0).sum(0), np.maximum(-out, 0).sum(0)
    pot0 = u0 * (pos + neg) - neg
    total = pot0 + lut[vid[-1]] - lut[vid[0]]
    assert np.allclose(pot + pe * pos - ne * neg, total, atol=1e-9)
    assert np.all((pot < pos) & (pot > -neg))


# ---- the C restatement (oracle/v2v_oracle_c.c) is pinned to the same golden vectors ----

@pytest.mark.parametrize("name", golden("esim").names("esim_"))
def test_c_oracle_esim(name):
    import v2v_oracle_c as orcc
    c = golden("esim").case(name)
    out = orcc.esim_video_to_voxel(c["video"], float(c["pos"]), float(c["neg"]), float(c["base_noise_std"]),
                                   c["u0"], c["hot"], c["g"], bool(c["external"]), lut=c["lut"])
    assert same(out, c["ref"])


@pytest.mark.parametrize("name", golden("scatter").names("scat_mv_"))
def test_c_oracle_make_voxel(name):
    import v2v_oracle_c as orcc
    c = golden("scatter").case(name)
    out = orcc.make_voxel(c["ts"], c["xs"], c["ys"], c["ps"], int(c["bins"]), int(c["H"]), int(c["W"]), bool(c["interp"]))
    assert same(out, c["ref"])


def test_c_and_numpy_oracles_agree_on_random_inputs():
    """The two restatements are independent (vectorised NumPy vs scalar C with numpy's divmod spelled out):
    they must agree bit for bit on random clips, thresholds incl. near-multiples, noise on/off/external."""
    import v2v_oracle_c as orcc
    from conftest import synth_video
    g = np.random.Generator(np.random.PCG64(99))
    lut = orc.esim_log_lut()
    for trial in range(12):
        n, h, w = int(g.integers(2, 9)), int(g.integers(1, 20)), int(g.integers(1, 20))
        vid = synth_video("iid" if trial % 2 else "walk", n, h, w, trial)
        pos = float(g.choice([0.05, 0.1, 0.2, 1 / 3, 0.7, 1.9]))
        neg = pos * float(g.choice([1.0, 1.25, 1.5, 1 / 1.5]))
        std = float(g.choice([0.0, 0.03, 0.1]))
        u0 = g.random((h, w))
        hot = np.where(g.random((h, w)) < 0.1, g.standard_normal((h, w)) * 5, 0.0)
        gs = g.standard_normal((n - 1, h, w))
        for ext in (False, True):
            a, pa = orc.esim_video_to_voxel(vid, pos, neg, std, u0, hot, gs, ext, lut, return_state=True)
            b, pb = orcc.esim_video_to_voxel(vid, pos, neg, std, u0, hot, gs, ext, lut, return_state=True)
            assert same(a, b) and same(pa, pb), (trial, ext)
    for trial in range(8):
        ne, h, w, bins = int(g.integers(1, 400)), int(g.integers(2, 30)), int(g.integers(2, 30)), int(g.choice([1, 3, 5, 15]))
        ts = np.sort(g.random(ne)) * float(g.choice([1e-3, 0.04, 2.0])) + float(g.choice([0.0, 17.3, 1.6e9]))
        if trial % 3 == 0:
            ts = ts.astype(np.float32)
        xs = g.integers(0, w, ne).astype(np.uint16)
        ys = g.integers(0, h, ne).astype(np.uint16)
        ps = (g.random(ne) < 0.5).astype(np.uint8)
        for interp in (False, True):
            assert same(orc.make_voxel(ts, xs, ys, ps, bins, h, w, interp), orcc.make_voxel(ts, xs, ys, ps, bins, h, w, interp)), (trial, interp)


def test_events_to_voxel_np():
    c = golden("scatter").case("scat_voxel_np")
    out = orc.events_to_voxel_np(c["xs"], c["ys"], c["ts"], c["ps"], int(c["bins"]), (int(c["H"]), int(c["W"])))
    assert same(out, c["ref"])


def test_rng_restatements_known_answers():
    """Philox4x32-10 against the Random123 known-answer vectors; xoshiro128++ against the reference C implementation's
    first outputs for the state {1,2,3,4}."""
    import v2v_oracle as orc
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    assert orc.xoshiro128pp([1, 2, 3, 4], 4) == [641, 1573767, 3222811527, 3517856514]


def test_direction_table_matches_library():
    """The oracle's restatement of the generator's direction table == the table compiled into the library (host copy
    through the C ABI; no GPU needed), and its entries are unit vectors to 21 bits."""
    import ctypes as C
    import v2v_oracle as orc
    from v2v_b200 import _lib
    buf = (C.c_uint32 * 4096)()
    _lib.check(_lib.load().v2v_noise_direction_table(buf))
    hi = np.frombuffer(buf, dtype=np.uint32).reshape(2048, 2).astype(np.uint64)
    lib_tab = (hi << np.uint64(32)).view(np.float64)
    tab = orc.esim_direction_table()
    assert np.array_equal(lib_tab, tab)
    assert np.abs(np.hypot(tab[:, 0], tab[:, 1]) - 1).max() < 2.0 ** -20
    # the 64-bit LCG behind the noise streams, against a direct big-integer evaluation
    w = orc.esim_noise_stream_words(123, 5, 77, 3)
    r = orc.philox4x32_10([77, 0, 5, 0x40000000], [123, 0])
    s, c = (r[1] << 32) | r[0], (r[3] << 32) | r[2] | 1
    s1 = (s * 0xF9B25D65 + c) % 2 ** 64
    s2 = (s1 * 0xF9B25D65 + c) % 2 ** 64
    assert w[:2] == [s1 >> 32, s2 >> 32]


def test_augment_oracle_vs_reference_goldens():
    """oracle/ restatement of data/esim_dataset.py:7-46,84-153 == the reference's recorded outputs (same seeds)."""
    import random
    import v2v_oracle as orc
    G = golden("augment")
    for name in G.names("noise_"):
        c = G.case(name)
        np.random.seed(int(c["seed"]))
        out = orc.add_noise_to_voxel(c["voxel"].copy(), float(c["noise_std"]), float(c["noise_fraction"]), bool(c["integer_noise"]))
        assert same(out, c["ref"])
    for name in G.names("hot_"):
        c = G.case(name)
        np.random.seed(int(c["np_seed"])), random.seed(int(c["py_seed"]))
        out = orc.add_hot_pixels_to_voxels(c["voxels"].copy(), float(c["hot_pixel_std"]), float(c["max_hot_pixel_fraction"]),
                                           bool(c["integer_noise"]))
        assert same(out, c["ref"])
    for name in G.names("item_"):
        c = G.case(name)
        np.random.seed(int(c["np_seed"])), random.seed(int(c["py_seed"]))
        fr, fl, vx, src = orc.cached_sequence_item(c["frames"], c["flow"], c["events"], int(c["sequence_length"]),
                                                   float(c["proba_pause_when_running"]), float(c["proba_pause_when_paused"]),
                                                   float(c["noise_std"]), float(c["noise_fraction"]), float(c["hot_pixel_std"]),
                                                   float(c["max_hot_pixel_fraction"]), bool(c["integer_noise"]))
        assert same(fr, c["ref_frame"]) and same(fl, c["ref_flow"]) and same(vx, c["ref_events"]) and same(src, c["src"])
