#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it (1) runs the reference function imported from /root/reference
(with the import stubs of SURVEY.md Appendix B — they back file I/O / plotting
only and are never called on the hot path), (2) asserts that the repo's CPU
oracle (oracle/v2v_oracle.py) is bit-identical on the same inputs and the same
``np.random`` stream, and (3) stores inputs, the pre-drawn random fields, the
LUT used, and the reference's outputs in a compact ``.npz``.

The GPU box has no /root/reference; its tests replay these files.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("V2V_REFERENCE", "/root/reference")

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import v2v_oracle as orc  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    for m in ("h5py", "matplotlib", "matplotlib.pyplot", "ffmpeg"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    evb = types.ModuleType("event_voxel_builder")
    evb.EventVoxelBuilder = object
    sys.modules["event_voxel_builder"] = evb
    import data.v2v_core_esim as esim
    import data.v2v_core_v2e as v2e
    import data.v2v_datasets as dsets
    import data.testh5 as testh5
    import utils.event_utils as eu
    return esim, v2e, dsets, testh5, eu


def synth_video(kind, n, h, w, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    if kind == "walk":
        base = g.integers(0, 256, size=(h, w)).astype(np.int64)
        steps = g.integers(-6, 7, size=(n, h, w))
        steps[0] = 0
        return np.clip(base[None] + np.cumsum(steps, axis=0), 0, 255).astype(np.uint8)
    if kind == "iid":
        return g.integers(0, 256, size=(n, h, w), dtype=np.uint8)
    if kind == "flash":            # all-black / all-white alternation with a static half
        v = np.zeros((n, h, w), dtype=np.uint8)
        v[1::2, :, : w // 2] = 255
        v[:, :, w // 2:] = 77
        return v
    if kind == "pause":            # walk with duplicated (paused) frames
        v = synth_video("walk", n, h, w, seed)
        idx = np.sort(g.integers(0, n, size=n))
        return v[idx]
    raise ValueError(kind)


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def gen_esim(esim, dsets, out):
    lut = orc.esim_log_lut()
    cases = [
        # name, kind, N,H,W, pos,neg, std, hot_frac, hot_std, external, seed
        ("walk_clean", "walk", 11, 24, 20, 0.2, 0.2, 0.0, 0.0, 0.0, False, 0),
        ("walk_noise", "walk", 11, 24, 20, 0.2, 0.2, 0.05, 0.05, 2.0, False, 1),
        ("walk_ext", "walk", 11, 24, 20, 0.3, 0.2, 0.05, 0.05, 2.0, True, 2),
        ("iid_small_thr", "iid", 11, 17, 33, 0.05, 0.075, 0.1, 0.001, 10.0, False, 3),
        ("iid_big_thr", "iid", 11, 17, 33, 1.3, 0.9, 0.02, 0.01, 0.1, False, 4),
        ("flash", "flash", 11, 16, 16, 0.2, 0.25, 0.0, 0.0, 0.0, False, 5),
        ("pause", "pause", 21, 12, 40, 0.11, 0.1, 0.01, 0.0, 0.0, False, 6),
        ("core_c1_shape", "walk", 40, 32, 32, 0.2, 0.2, 0.0, 0.0, 0.0, False, 7),
        ("ragged", "iid", 6, 7, 13, 0.4, 0.6, 0.1, 0.2, 1.0, False, 8),
        ("ragged_ext", "iid", 6, 5, 3, 0.4, 0.6, 0.1, 0.2, 1.0, True, 9),
    ]
    for (name, kind, n, h, w, pos, neg, std, hf, hs, ext, seed) in cases:
        video = synth_video(kind, n, h, w, 100 + seed)
        np.random.seed(seed)
        ref = esim.EventEmulator(pos_thres=pos, neg_thres=neg, base_noise_std=std, hot_pixel_fraction=hf,
                                 hot_pixel_std=hs, put_noise_external=ext).video_to_voxel(video)
        np.random.seed(seed)
        u0, hot, g = orc.esim_draw_randomness(n, h, w, hf, hs)
        mine = orc.esim_video_to_voxel(video, pos, neg, std, u0, hot, g, ext, lut=lut)
        assert same(ref, mine), f"oracle != reference for esim case {name}"
        # LUT identity against the reference's own log image
        ref_log = np.log(0.001 + esim.reverse_gamma_correction(video) / 255.0)
        assert same(ref_log, lut[video]), name
        out[f"esim_{name}"] = dict(video=video, pos=pos, neg=neg, base_noise_std=std, hot_pixel_fraction=hf,
                                   hot_pixel_std=hs, external=int(ext), seed=seed, u0=u0, hot=hot, g=g,
                                   lut=lut, ref=ref)

    # the dataset-level entry: parameter sampling + binning (data/v2v_datasets.py:363-410)
    ds_cases = [
        ("ds_default", dict(), 5, 1, 11, 16, 24, 11),
        ("ds_train_cfg", dict(base_noise_std_range=[0, 0.1], hot_pixel_std_range=[0, 10]), 5, 1, 11, 16, 24, 12),
        ("ds_fpb2", dict(), 5, 2, 21, 10, 12, 13),
        ("ds_ext_scaled", dict(put_noise_external=True, scale_noise_strength=True), 5, 1, 11, 8, 8, 14),
        ("ds_scaled", dict(scale_noise_strength=True), 3, 2, 13, 8, 8, 15),
        ("ds_fixed_thr", dict(use_fixed_thresholds=True), 5, 1, 11, 8, 16, 16),
    ]
    for (name, cfg, bins, fpb, n, h, w, seed) in ds_cases:
        ds = dsets.WebvidDatasetV2.__new__(dsets.WebvidDatasetV2)
        ds.load_configs(dict(cfg))
        video = synth_video("walk", n, h, w, 200 + seed)
        fixed = (0.31, 0.27) if cfg.get("use_fixed_thresholds") else (None, None)
        np.random.seed(seed)
        params, ref = ds.imgs_to_voxels(video, bins, fpb, 24, fixed[0], fixed[1])
        np.random.seed(seed)
        p2 = orc.sample_esim_params(
            np.random, ds.threshold_range, ds.max_thres_pos_neg_gap, ds.base_noise_std_range,
            ds.hot_pixel_fraction_range, ds.hot_pixel_std_range, ds.scale_noise_strength,
            ds.put_noise_external, fixed[0], fixed[1])
        assert p2 == params, (name, p2, params)
        u0, hot, g = orc.esim_draw_randomness(n, h, w, p2["hot_pixel_fraction"], p2["hot_pixel_std"])
        iv = orc.esim_video_to_voxel(video, p2["pos_thres"], p2["neg_thres"], p2["base_noise_std"], u0, hot, g,
                                     ds.put_noise_external, lut=lut)
        mine = orc.bin_accumulate(iv, bins, fpb)
        assert same(ref, mine), f"oracle != reference for dataset case {name}"
        out[f"esimds_{name}"] = dict(video=video, bins=bins, fpb=fpb, seed=seed, external=int(ds.put_noise_external),
                                     u0=u0, hot=hot, g=g, lut=lut, ref=ref,
                                     cfg_keys=np.array(sorted(cfg.keys()), dtype="U64"),
                                     cfg_vals=np.array([repr(cfg[k]) for k in sorted(cfg.keys())], dtype="U64"),
                                     fixed_pos=-1.0 if fixed[0] is None else fixed[0],
                                     fixed_neg=-1.0 if fixed[1] is None else fixed[1],
                                     **{f"p_{k}": v for k, v in params.items()})

    # frame packing (data/v2v_datasets.py:328-338,352) — restated, checked against torch
    import torch
    imgs = synth_video("iid", 11, 6, 10, 999)[..., None]
    for add in (False, True):
        fpi, cnt = 5, 2
        if not add:
            ref = torch.stack([torch.tensor(imgs[(i + 1) * fpi].copy(), dtype=torch.float32).permute(2, 0, 1)
                               for i in range(cnt)], axis=0) / 255
        else:
            ref = torch.stack([torch.tensor(imgs[i * fpi].copy(), dtype=torch.float32).permute(2, 0, 1)
                               for i in range(cnt + 1)], axis=0) / 255
        mine = orc.pack_frames(imgs, fpi, cnt, add)
        assert same(ref.numpy(), mine)
        out[f"frames_add{int(add)}"] = dict(imgs=imgs, fpi=fpi, cnt=cnt, add=int(add), ref=ref.numpy())


V2E_PRESETS = {
    # presets of data/v2v_core_v2e.py:354-375 mapped on the current ctor
    "clean": dict(threshold_model="pn_related", thres_mean_mean=0.2, thres_mean_std=0.02, thres_diff_mean=0.0,
                  thres_diff_std=0.02, cutoff_hz=0.0, leak_rate_hz=0.0, shot_noise_rate_hz=0.0,
                  leak_jitter_fraction=0.0, noise_rate_cov_decades=0.0),
    "noisy": dict(threshold_model="pn_related", thres_mean_mean=0.2, thres_mean_std=0.05, thres_diff_mean=0.0,
                  thres_diff_std=0.05, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0,
                  leak_jitter_fraction=0.1, noise_rate_cov_decades=0.1),
    "leak_only": dict(threshold_model="pn_related", thres_mean_mean=0.3, thres_mean_std=0.05, thres_diff_mean=0.02,
                      thres_diff_std=0.03, cutoff_hz=0.0, leak_rate_hz=2.0, shot_noise_rate_hz=0.0,
                      leak_jitter_fraction=0.3, noise_rate_cov_decades=0.2),
    "shot_only": dict(threshold_model="spatial_independent", thres_mean_mean=0.25, thres_mean_std=0.04,
                      thres_diff_mean=0.0, thres_diff_std=0.0, cutoff_hz=0.0, leak_rate_hz=0.0,
                      shot_noise_rate_hz=40.0, leak_jitter_fraction=0.0, noise_rate_cov_decades=0.0),
    "cutoff_only": dict(threshold_model="pn_related", thres_mean_mean=0.15, thres_mean_std=0.03, thres_diff_mean=0.0,
                        thres_diff_std=0.03, cutoff_hz=15.0, leak_rate_hz=0.0, shot_noise_rate_hz=0.0,
                        leak_jitter_fraction=0.0, noise_rate_cov_decades=0.0),
    # the per-frame threshold model (data/v2v_core_v2e.py:417-421): both maps re-drawn before every frame
    "perframe_clean": dict(threshold_model="spatial_temporal_independent", thres_mean_mean=0.2, thres_mean_std=0.03,
                           thres_diff_mean=0.0, thres_diff_std=0.0, cutoff_hz=0.0, leak_rate_hz=0.0, shot_noise_rate_hz=0.0,
                           leak_jitter_fraction=0.0, noise_rate_cov_decades=0.0),
    "perframe_noisy": dict(threshold_model="spatial_temporal_independent", thres_mean_mean=0.2, thres_mean_std=0.05,
                           thres_diff_mean=0.0, thres_diff_std=0.0, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0,
                           leak_jitter_fraction=0.1, noise_rate_cov_decades=0.1),
}


def hdr_degrade(video, scale):
    # data/v2v_datasets.py:473-477
    return np.clip((video - 127.5) * scale + 127.5, 0, 255).astype(np.uint8)


def gen_v2e(v2e, out):
    lut = orc.v2e_log_lut()
    k = 0
    for preset, params in V2E_PRESETS.items():
        kinds = [("walk", 11, 12, 16), ("iid", 6, 9, 7)]
        if preset in ("noisy", "cutoff_only", "shot_only"):
            kinds.append(("u8", 9, 10, 12))
        for kind, n, h, w in kinds:
            seed = 40 + k if kind != "u8" else 140 + k
            k += kind != "u8"
            video8 = hdr_degrade(synth_video("walk" if kind == "u8" else kind, n, h, w, 300 + seed), 2.1)
            video = video8.astype(np.float64)
            if kind == "u8":
                # uint8 handed to the reference as is: rescale_intensity_frame (:190) wraps for values >= 236
                video = video8
                assert (video8 >= 236).mean() > 0.02
            ref = v2e.video_to_voxel(video, 24, refractory_period_s=0, seed=seed, **params)
            rec = {}
            np.random.seed(seed)
            mine = orc.v2e_video_to_voxel(video, 24, params, np.random, lut=lut, record=rec)
            assert same(ref, mine), f"oracle != reference for v2e {preset}/{kind}"
            if params["threshold_model"] != "spatial_temporal_independent":
                again = orc.v2e_replay(video, 24, params, rec, lut=lut)
                assert same(ref, again)
            # lin_log LUT identity
            assert same(v2e.lin_log(video[0]), lut[video8[0]])
            d = dict(video=video8, fps=24, seed=seed, lut=lut, ref=ref, u8_input=int(kind == "u8"),
                     thr_a=rec["thr_a"], thr_b=rec["thr_b"], noise_randn=rec["noise_randn"],
                     pos_thres=rec["pos_thres"], neg_thres=rec["neg_thres"], noise_rate=rec["noise_rate"],
                     threshold_model=np.array(params["threshold_model"]))
            for key, val in params.items():
                if key != "threshold_model":
                    d[f"p_{key}"] = val
            if rec["pos_thres_frames"]:
                d["pos_thres_frames"] = np.stack(rec["pos_thres_frames"])
                d["neg_thres_frames"] = np.stack(rec["neg_thres_frames"])
            if rec["leak_randn"]:
                d["leak_randn"] = np.stack(rec["leak_randn"])
            if rec["pos_shot"]:
                d["pos_shot"] = np.stack(rec["pos_shot"]).astype(np.int32)
                d["neg_shot"] = np.stack(rec["neg_shot"]).astype(np.int32)
            out[f"v2e_{preset}_{kind}"] = d


def synth_events(ne, h, w, seed, ts_dtype=np.float64, span=0.04, t0=12.5, hot=False):
    g = np.random.Generator(np.random.PCG64(seed))
    xs = g.integers(0, w, size=ne).astype(np.uint16)
    ys = g.integers(0, h, size=ne).astype(np.uint16)
    if hot and ne:
        sel = g.random(ne) < 0.3
        xs[sel] = w - 1
        ys[sel] = h - 1
    ts = np.sort(g.random(ne) * span + t0).astype(ts_dtype)
    ps = (g.random(ne) < 0.5).astype(np.uint8)
    return ts, xs, ys, ps


def gen_scatter(testh5, eu, out):
    import torch
    # --- test-loop flavour: TestH5Dataset.make_voxel ---
    cases = [
        ("mv_disc5", 3000, 26, 35, 5, False, np.float64, False),
        ("mv_interp5", 3000, 26, 35, 5, True, np.float64, False),
        ("mv_disc15", 2000, 20, 24, 15, False, np.float64, True),
        ("mv_interp15", 2000, 20, 24, 15, True, np.float64, True),
        ("mv_disc5_f32ts", 1500, 18, 22, 5, False, np.float32, False),
        ("mv_interp5_f32ts", 1500, 18, 22, 5, True, np.float32, False),
        ("mv_empty", 0, 8, 9, 5, False, np.float64, False),
        ("mv_single", 1, 8, 9, 5, True, np.float64, False),
        ("mv_single_disc", 1, 8, 9, 5, False, np.float64, False),
    ]
    for i, (name, ne, h, w, bins, interp, tdt, hot) in enumerate(cases):
        ts, xs, ys, ps = synth_events(ne, h, w, 500 + i, tdt, hot=hot)
        ds = testh5.TestH5Dataset.__new__(testh5.TestH5Dataset)
        ds.num_bins, ds.H, ds.W, ds.interpolate_bins = bins, h, w, interp
        ref = ds.make_voxel([ts, xs, ys, ps])
        mine = orc.make_voxel(ts, xs, ys, ps, bins, h, w, interp)
        assert same(ref, mine), name
        out[f"scat_{name}"] = dict(ts=ts, xs=xs, ys=ys, ps=ps, bins=bins, H=h, W=w, interp=int(interp), ref=ref)
    # same-timestamp window (τ_last = 0)
    ts, xs, ys, ps = synth_events(50, 8, 9, 577)
    ts[:] = ts[0]
    for interp in (False, True):
        ds = testh5.TestH5Dataset.__new__(testh5.TestH5Dataset)
        ds.num_bins, ds.H, ds.W, ds.interpolate_bins = 5, 8, 9, interp
        ref = ds.make_voxel([ts, xs, ys, ps])
        assert same(ref, orc.make_voxel(ts, xs, ys, ps, 5, 8, 9, interp))
        out[f"scat_mv_samets_{int(interp)}"] = dict(ts=ts, xs=xs, ys=ys, ps=ps, bins=5, H=8, W=9,
                                                    interp=int(interp), ref=ref)

    # --- legacy torch flavour: events_to_voxel_torch & friends ---
    tcases = [("tv_bil5", 4000, 18, 24, 5, True), ("tv_disc5", 4000, 18, 24, 5, False),
              ("tv_bil3", 500, 10, 12, 3, True), ("tv_disc9", 700, 10, 12, 9, False)]
    for i, (name, ne, h, w, bins, bil) in enumerate(tcases):
        ts, xs, ys, ps = synth_events(ne, h, w, 600 + i, np.float64, span=0.5, t0=0.0, hot=(i % 2 == 1))
        xt = torch.from_numpy(xs.astype(np.float32))
        yt = torch.from_numpy(ys.astype(np.float32))
        tt = torch.from_numpy((ts - ts[0]).astype(np.float32))
        pt = torch.from_numpy(ps.astype(np.float32) * 2 - 1)
        ref = eu.events_to_voxel_torch(xt, yt, tt, pt, bins, sensor_size=(h, w), temporal_bilinear=bil).numpy()
        mine = orc.events_to_voxel_f32(xt.numpy(), yt.numpy(), tt.numpy(), pt.numpy(), bins, (h, w), bil)
        assert same(ref, mine), name
        rp, rn = eu.events_to_neg_pos_voxel_torch(xt, yt, tt, pt, bins, sensor_size=(h, w), temporal_bilinear=bil)
        mp, mn = orc.events_to_neg_pos_voxel_f32(xt.numpy(), yt.numpy(), tt.numpy(), pt.numpy(), bins, (h, w), bil)
        assert same(rp.numpy(), mp) and same(rn.numpy(), mn), name
        out[f"scat_{name}"] = dict(xs=xt.numpy(), ys=yt.numpy(), ts=tt.numpy(), ps=pt.numpy(), bins=bins, H=h, W=w,
                                   bilinear=int(bil), ref=ref, ref_pos=rp.numpy(), ref_neg=rn.numpy())

    # --- event images ---
    g = np.random.Generator(np.random.PCG64(700))
    h, w, ne = 14, 19, 1200
    xf = (g.random(ne) * (w + 1.5)).astype(np.float32)          # some out of range on purpose
    yf = (g.random(ne) * (h + 1.5)).astype(np.float32)
    pf = (g.random(ne) * 2 - 1).astype(np.float32)
    for name, kw in (("img_bil_pad", dict(interpolation="bilinear", padding=True, clip_out_of_range=True)),
                     ("img_nearest_clip_nopad", dict(interpolation=None, padding=False, clip_out_of_range=True))):
        # the nearest branch never applies the clip mask (utils/event_utils.py:371-375), so
        # out-of-range events raise IndexError there: keep that case in range
        xin, yin = (xf, yf) if kw["interpolation"] else (np.minimum(xf, w - 1), np.minimum(yf, h - 1))
        ref = eu.events_to_image_torch(torch.from_numpy(xin), torch.from_numpy(yin), torch.from_numpy(pf),
                                       sensor_size=(h, w), **kw).numpy()
        mine = orc.events_to_image_f32(xin, yin, pf, sensor_size=(h, w), **kw)
        assert same(ref, mine), name
        out[f"scat_{name}"] = dict(xs=xin, ys=yin, ps=pf, H=h, W=w, ref=ref,
                                   bilinear=int(kw["interpolation"] == "bilinear"), padding=int(kw["padding"]),
                                   clip=int(kw["clip_out_of_range"]))
    # in-range nearest with the default padding=True & clip (clip bound = size-1)
    xi = g.integers(0, w, size=ne).astype(np.float32)
    yi = g.integers(0, h, size=ne).astype(np.float32)
    ref = eu.events_to_image_torch(torch.from_numpy(xi), torch.from_numpy(yi), torch.from_numpy(pf),
                                   sensor_size=(h, w)).numpy()
    mine = orc.events_to_image_f32(xi, yi, pf, sensor_size=(h, w))
    assert same(ref, mine)
    out["scat_img_nearest_default"] = dict(xs=xi, ys=yi, ps=pf, H=h, W=w, ref=ref, bilinear=0, padding=1, clip=1)
    # numpy bincount image
    ref = eu.events_to_image(xi.astype(np.int64), yi.astype(np.int64), pf.astype(np.float64), sensor_size=(h, w))
    mine = orc.events_to_image_np(xi.astype(np.int64), yi.astype(np.int64), pf.astype(np.float64), (h, w))
    assert same(ref, mine)
    # numpy events_to_voxel (bilinear branch; needs ts/ps as [N,1] columns)
    ne2 = 900
    xs2, ys2 = g.integers(0, w, ne2), g.integers(0, h, ne2)
    ts2 = np.sort(g.random(ne2)) * 0.3 + 2.0
    ps2 = (g.integers(0, 2, ne2) * 2 - 1).astype(np.float64)
    ref_v = eu.events_to_voxel(xs2, ys2, ts2[:, None], ps2[:, None], 5, sensor_size=(h, w), temporal_bilinear=True)
    assert same(ref_v, orc.events_to_voxel_np(xs2, ys2, ts2, ps2, 5, (h, w)))
    out["scat_voxel_np"] = dict(xs=xs2, ys=ys2, ts=ts2, ps=ps2, bins=5, H=h, W=w, ref=ref_v)
    out["scat_img_np"] = dict(xs=xi.astype(np.int64), ys=yi.astype(np.int64), ps=pf.astype(np.float64), H=h, W=w,
                              ref=ref)


def gen_frames(dsets, out):
    """Frame-side packing: pause gather, HDR/LDR degrade, bgr_to_gray — outputs of the reference's own code."""
    import textwrap
    # -- pause sequence: the reference has it inline in __getitem__; execute exactly those source lines
    lines = open(os.path.join(REF, "data", "v2v_datasets.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "img_idxes = []")
    i1 = next(i for i, l in enumerate(lines) if l.strip() == "true_img_cnt = idx + 1")
    block = textwrap.dedent("\n".join(lines[i0:i1 + 1]).replace("\t", "    "))
    for ci, (seed, p_run, p_paused, img_cnt, fpi, add_evs) in enumerate([
            (0, 0.0102, 0.9791, 24, 5, False), (1, 0.05, 0.9, 24, 5, False), (2, 0.3, 0.5, 10, 10, True),
            (3, 0.0, 0.9, 8, 5, False), (4, 1.0, 1.0, 6, 5, False), (5, 0.5, 0.0, 12, 5, True)]):
        fake = types.SimpleNamespace(proba_pause_when_running=p_run, proba_pause_when_paused=p_paused, frames_per_img=fpi,
                                     output_additional_evs=add_evs)
        ns = {"np": np, "self": fake, "start_frame": 7, "img_cnt": img_cnt}
        np.random.seed(seed)
        exec(block, ns)
        count = img_cnt * fpi + 1 + (fpi if add_evs else 0)
        np.random.seed(seed)
        o_idx, o_cnt = orc.pause_indices(count, p_run, p_paused)
        assert list(o_idx) == list(ns["img_idxes"]) and o_cnt == ns["true_img_cnt"]
        out[f"pause_{ci}"] = dict(seed=seed, p_run=p_run, p_paused=p_paused, count=count,
                                  img_idxes=np.asarray(ns["img_idxes"], dtype=np.int32), true_img_cnt=ns["true_img_cnt"])
    # -- degrade_video (hdr / ldr)
    for ci, (kind, seed) in enumerate([("hdr", 0), ("hdr", 1), ("ldr", 2), ("ldr", 3)]):
        g = np.random.Generator(np.random.PCG64(seed))
        imgs = [g.integers(0, 256, (12, 16, 1), dtype=np.uint8) for _ in range(5)]
        imgs[0][:, :, 0].flat[:256] = np.arange(256, dtype=np.uint8)          # every value occurs
        fake = types.SimpleNamespace(video_degrade=kind)
        np.random.seed(seed)
        ref = dsets.WebvidDatasetV2.degrade_video(fake, [im.copy() for im in imgs])
        np.random.seed(seed)
        mine, scale = orc.degrade_video([im.copy() for im in imgs], kind)
        assert all(same(a, b) for a, b in zip(ref, mine))
        out[f"degrade_{ci}"] = dict(kind=np.array(kind), seed=seed, scale=scale, imgs=np.stack(imgs), ref=np.stack(ref))
    # -- bgr_to_gray: every colour triple whose exact weighted sum is an integer (the summation order decides the result) + random ones
    a, b, c = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    sens = (5870 * a + 1140 * b + 2989 * c) % 10000 == 0
    tri = np.stack([a[sens], b[sens], c[sens]], -1).astype(np.uint8)
    g = np.random.Generator(np.random.PCG64(9))
    tri = np.concatenate([tri, g.integers(0, 256, (20000, 3), dtype=np.uint8)])
    img = np.concatenate([tri, g.integers(0, 256, (tri.shape[0], 1), dtype=np.uint8)], -1)[None]     # [1, n, 4]: BGRA-like, 4th channel ignored
    ref = dsets.bgr_to_gray(img)
    assert same(ref, orc.bgr_to_gray(img))
    out["bgr_to_gray"] = dict(img=img, ref=ref, n_sensitive=int(sens.sum()))


def gen_augment(out):
    """Voxel-space augmentation of the cached-voxel dataset: the reference's own add_noise_to_voxel /
    add_hot_pixels_to_voxels / ESIMH5Dataset.__getitem__ (data/esim_dataset.py:7-46,84-153).  The dataset object is
    built without its constructor and reads from a dict of arrays instead of an h5 file (the I/O is out of scope)."""
    import random
    import data.esim_dataset as ed
    g = np.random.Generator(np.random.PCG64(99))
    vox = g.integers(-3, 4, size=(6, 5, 24, 40)).astype(np.float32)
    k = 0
    for integer_noise in (False, True):
        for frac in (0.1, 1.0):
            np.random.seed(40 + k)
            ref = ed.add_noise_to_voxel(vox[0].copy(), 0.7, frac, integer_noise)
            np.random.seed(40 + k)
            mine = orc.add_noise_to_voxel(vox[0].copy(), 0.7, frac, integer_noise)
            assert same(ref, mine), "oracle != reference for add_noise_to_voxel"
            out[f"noise_int{int(integer_noise)}_frac{frac}"] = dict(voxel=vox[0], noise_std=0.7, noise_fraction=frac,
                                                                    integer_noise=int(integer_noise), seed=40 + k, ref=ref)
            k += 1
        np.random.seed(60 + k), random.seed(7 + k)
        ref = ed.add_hot_pixels_to_voxels(vox.astype(np.float32).copy(), 2.0, 0.05, integer_noise)
        np.random.seed(60 + k), random.seed(7 + k)
        mine = orc.add_hot_pixels_to_voxels(vox.astype(np.float32).copy(), 2.0, 0.05, integer_noise)
        assert same(ref, mine) and (ref != vox).any(), "oracle != reference for add_hot_pixels_to_voxels"
        out[f"hot_int{int(integer_noise)}"] = dict(voxels=vox, hot_pixel_std=2.0, max_hot_pixel_fraction=0.05,
                                                    integer_noise=int(integer_noise), np_seed=60 + k, py_seed=7 + k, ref=ref)
    # the dataset item: crop and flip off (views), pause + noise + hot pixels as the reference runs them
    S, T, Hh, Ww = 14, 10, 20, 28
    frames = g.random((S, 1, Hh, Ww)).astype(np.float32)
    flow = g.standard_normal((S, 2, Hh, Ww)).astype(np.float32)
    events = g.integers(-4, 5, size=(S, 5, Hh, Ww)).astype(np.float32)
    for case, (integer_noise, frac, pr, pp) in enumerate([(False, 1.0, 0.05, 0.9), (False, 0.3, 0.4, 0.6), (True, 1.0, 0.3, 0.9)]):
        ds = ed.ESIMH5Dataset.__new__(ed.ESIMH5Dataset)
        ds.h5_file = {"frames": frames, "flow": flow, "events": events}
        ds.sequence_length, ds.samples = T, [(2, 2 + T)]
        ds.proba_pause_when_running, ds.proba_pause_when_paused = pr, pp
        ds.noise_std, ds.noise_fraction, ds.hot_pixel_std, ds.max_hot_pixel_fraction = 0.6, frac, 1.5, 0.03
        ds.random_crop_size, ds.random_flip, ds.integer_noise, ds.data_source_idx = None, False, integer_noise, 0
        np.random.seed(80 + case), random.seed(3 + case)
        item = ds[0]
        np.random.seed(80 + case), random.seed(3 + case)
        fr, fl, vx, src = orc.cached_sequence_item(frames[2:2 + T], flow[2:2 + T], events[2:2 + T], T, pr, pp, 0.6, frac, 1.5,
                                                   0.03, integer_noise)
        assert same(item["frame"].numpy(), fr) and same(item["flow"].numpy(), fl) and same(item["events"].numpy(), vx), \
            "oracle != reference for ESIMH5Dataset.__getitem__"
        assert (src < 0).any() or case == 0
        out[f"item_{case}"] = dict(frames=frames[2:2 + T], flow=flow[2:2 + T], events=events[2:2 + T], sequence_length=T,
                                   proba_pause_when_running=pr, proba_pause_when_paused=pp, noise_std=0.6, noise_fraction=frac,
                                   hot_pixel_std=1.5, max_hot_pixel_fraction=0.03, integer_noise=int(integer_noise),
                                   np_seed=80 + case, py_seed=3 + case, ref_frame=fr, ref_flow=fl, ref_events=vx, src=src)


def main():
    esim, v2e, dsets, testh5, eu = import_reference()
    only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
    groups = {"esim": {}, "v2e": {}, "scatter": {}, "frames": {}, "augment": {}}
    if only:
        groups = {k: v for k, v in groups.items() if k in only}
    if "esim" in groups:
        gen_esim(esim, dsets, groups["esim"])
    if "v2e" in groups:
        gen_v2e(v2e, groups["v2e"])
    if "scatter" in groups:
        gen_scatter(testh5, eu, groups["scatter"])
    if "frames" in groups:
        gen_frames(dsets, groups["frames"])
    if "augment" in groups:
        gen_augment(groups["augment"])
    import torch
    meta = dict(numpy=np.__version__, torch=torch.__version__)
    for gname, cases in groups.items():
        flat = {"__numpy__": np.array(meta["numpy"]), "__torch__": np.array(meta["torch"]),
                "__cases__": np.array(sorted(cases.keys()), dtype="U64")}
        for cname, d in cases.items():
            for k, v in d.items():
                flat[f"{cname}/{k}"] = np.asarray(v)
        path = os.path.join(HERE, f"{gname}.npz")
        np.savez_compressed(path, **flat)
        print(f"{path}: {len(cases)} cases, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
