"""GPU parity of the ESIM frame->voxel kernel against the reference's golden outputs and the CPU oracle.

Bar: crossing counts bit-exact; voxels bit-exact as float32 (integer valued) and within 1e-5 relative
in the external-noise mode (float64 sums rounded to float32).
"""
import os

import numpy as np
import pytest
import torch

import v2v_oracle as orc
from conftest import golden, synth_video
from v2v_b200 import _lib

pytestmark = pytest.mark.gpu


def run_explicit(v2v, dev, video, pos, neg, std, u0, hot, g, external, lut, bins=1, fpb=1, **kw):
    fr = torch.from_numpy(video).to(dev)
    return v2v.frames_to_voxel(fr, pos, neg, num_bins=bins, frames_per_bin=fpb, noise="explicit",
                               base_noise_std=std, put_noise_external=external, u0=u0[None], hot_noise=hot[None],
                               base_gauss=g[None], lut=lut, **kw)


@pytest.mark.parametrize("name", golden("esim").names("esim_"))
def test_golden_core(cuda_device, name):
    import v2v_b200 as v2v
    c = golden("esim").case(name)
    o = run_explicit(v2v, cuda_device, c["video"], float(c["pos"]), float(c["neg"]), float(c["base_noise_std"]),
                     c["u0"], c["hot"], c["g"], bool(c["external"]), c["lut"], with_stats=True)
    got = o.voxel[0, :, 0].cpu().numpy()
    ref = c["ref"]
    if bool(c["external"]):
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)
        assert np.array_equal(got, ref.astype(np.float32))       # in fact the float32 rounding of the reference
    else:
        assert np.array_equal(got.astype(np.float64), ref)       # bit-exact counts
        st = o.stats[0].cpu().numpy()
        assert st[0] == int(np.maximum(ref, 0).sum()) and st[1] == int(np.maximum(-ref, 0).sum())


@pytest.mark.parametrize("name", golden("esim").names("esimds_"))
def test_golden_dataset_level(cuda_device, name):
    import v2v_b200 as v2v
    c = golden("esim").case(name)
    o = run_explicit(v2v, cuda_device, c["video"], float(c["p_pos_thres"]), float(c["p_neg_thres"]),
                     float(c["p_base_noise_std"]), c["u0"], c["hot"], c["g"], bool(c["external"]), c["lut"],
                     bins=int(c["bins"]), fpb=int(c["fpb"]))
    got = o.voxel[0].cpu().numpy()
    assert got.shape == c["ref"].shape
    assert np.allclose(got, c["ref"], rtol=1e-5, atol=1e-6)
    if not bool(c["external"]):
        assert np.array_equal(got.astype(np.float64), c["ref"])


@pytest.mark.parametrize("name", golden("esim").names("esim_"))
def test_event_emulator_numpy_rng_is_reference(cuda_device, name):
    """Reference signature + same np.random.seed -> the reference's output (replayed MT19937 stream)."""
    import v2v_b200 as v2v
    c = golden("esim").case(name)
    np.random.seed(int(c["seed"]))
    em = v2v.EventEmulator(pos_thres=float(c["pos"]), neg_thres=float(c["neg"]), base_noise_std=float(c["base_noise_std"]),
                           hot_pixel_fraction=float(c["hot_pixel_fraction"]), hot_pixel_std=float(c["hot_pixel_std"]),
                           put_noise_external=bool(c["external"]), rng="numpy")
    got = em.video_to_voxel(c["video"], lut=c["lut"])
    assert got.dtype == np.float64 and got.shape == c["ref"].shape
    assert np.allclose(got, c["ref"], rtol=1e-5, atol=1e-6)
    if not bool(c["external"]):
        assert np.array_equal(got, c["ref"])


@pytest.mark.parametrize("name", golden("esim").names("esimds_"))
def test_imgs_to_voxels_mixin(cuda_device, name):
    import v2v_b200 as v2v
    c = golden("esim").case(name)
    cfg = {str(k): eval(str(v)) for k, v in zip(c["cfg_keys"], c["cfg_vals"])}
    vz = v2v.V2VVoxelizer(cfg, rng="numpy")
    fixed = (None, None) if float(c["fixed_pos"]) < 0 else (float(c["fixed_pos"]), float(c["fixed_neg"]))
    import v2v_b200.esim as E
    old = E.esim_log_lut
    E._lut_cache.clear()
    E.esim_log_lut = lambda: c["lut"]            # replay the LUT the fixture was generated with
    try:
        np.random.seed(int(c["seed"]))
        params, vox = vz.imgs_to_voxels(c["video"], int(c["bins"]), int(c["fpb"]), 24, fixed[0], fixed[1])
    finally:
        E.esim_log_lut = old
        E._lut_cache.clear()
    for k in ("pos_thres", "neg_thres", "base_noise_std", "hot_pixel_fraction", "hot_pixel_std"):
        assert params[k] == float(c[f"p_{k}"])
    assert np.allclose(vox, c["ref"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape,kind,pos,neg", [
    ((41, 256, 256), "walk", 0.2, 0.2),        # BASELINE config 1 shape through the binned path
    ((40, 256, 256), "iid", 0.2, 0.2),         # config 1 raw core, stress content
    ((11, 96, 160), "iid", 0.05, 0.0625),      # smallest shipped thresholds, many multi-count crossings
    ((11, 480, 640), "walk", 0.7, 1.05),       # config 2 frame size (vectorised path)
    ((6, 33, 17), "iid", 0.31, 0.2),           # ragged plane: HW % 4 != 0 (scalar path)
])
def test_oracle_parity_sizes(cuda_device, shape, kind, pos, neg):
    import v2v_b200 as v2v
    n, h, w = shape
    vid = synth_video(kind, n, h, w, 42)
    g = np.random.Generator(np.random.PCG64(7))
    u0, hot = g.random((h, w)), np.where(g.random((h, w)) < 0.01, g.standard_normal((h, w)) * 3.0, 0.0)
    gs = g.standard_normal((n - 1, h, w))
    lut = orc.esim_log_lut()
    for ext in (False, True):
        ref = orc.esim_video_to_voxel(vid, pos, neg, 0.04, u0, hot, gs, ext, lut=lut)
        o = run_explicit(v2v, cuda_device, vid, pos, neg, 0.04, u0, hot, gs, ext, lut, return_potential=True)
        got = o.voxel[0, :, 0].cpu().numpy()
        if ext:
            assert np.array_equal(got, ref.astype(np.float32))
        else:
            assert np.array_equal(got.astype(np.float64), ref)


def test_near_multiple_thresholds(cuda_device):
    """Adversarial floor-division cases: potentials that are k*thr +- a few ulp."""
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    h, w = 64, 64
    g = np.random.Generator(np.random.PCG64(11))
    vid = np.zeros((2, h, w), dtype=np.uint8)
    vid[0] = g.integers(0, 256, (h, w))
    vid[1] = g.integers(0, 256, (h, w))
    for thr in (0.05, 0.1, 0.2, 0.3, 1.0 / 3.0, 0.7):
        d = lut[vid[1]] - lut[vid[0]]
        k = g.integers(1, 40, (h, w)) * np.sign(g.standard_normal((h, w)))
        target = k * thr
        for _ in range(int(g.integers(0, 4))):
            target = np.nextafter(target, np.inf if g.random() < 0.5 else -np.inf)
        pot0 = target - d                       # potential lands (almost) on a multiple of the threshold
        # feed pot0 through potential_in (u0 path would re-derive it)
        fr = torch.from_numpy(vid).to(cuda_device)
        o = v2v.frames_to_voxel(fr, thr, thr, num_bins=1, potential_in=pot0[None], return_potential=True, lut=lut)
        got = o.voxel[0, 0, 0].cpu().numpy().astype(np.float64)
        pot = pot0 + d
        pe = np.where(pot >= thr, np.floor_divide(pot, thr), 0)
        ne = np.where(pot <= -thr, np.floor_divide(-pot, thr), 0)
        assert np.array_equal(got, pe - ne)
        exp_pot = pot - pe * thr
        exp_pot = exp_pot + ne * thr
        assert np.array_equal(o.potential[0].cpu().numpy(), exp_pot)


def test_batched_per_clip_and_per_pixel_thresholds(cuda_device):
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    B, n, h, w = 3, 11, 40, 64
    vids = np.stack([synth_video("walk", n, h, w, 60 + b) for b in range(B)])
    g = np.random.Generator(np.random.PCG64(5))
    u0 = g.random((B, h, w))
    pos = np.array([0.1, 0.25, 0.8])
    neg = np.array([0.15, 0.2, 0.6])
    fr = torch.from_numpy(vids).to(cuda_device)
    o = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, u0=u0, lut=lut, with_stats=True)
    z, zg = np.zeros((h, w)), np.zeros((n - 1, h, w))
    for b in range(B):
        ref = orc.bin_accumulate(orc.esim_video_to_voxel(vids[b], pos[b], neg[b], 0.0, u0[b], z, zg, False, lut), 5, 1)
        assert np.array_equal(o.voxel[b].cpu().numpy().astype(np.float64), ref)
    # per-pixel maps: each pixel must behave like a clip with that scalar threshold
    pmap = g.uniform(0.05, 1.0, (B, h, w))
    nmap = g.uniform(0.05, 1.0, (B, h, w))
    o2 = v2v.frames_to_voxel(fr, pmap, nmap, num_bins=5, u0=u0, lut=lut)
    got = o2.voxel.cpu().numpy()
    ys, xs = g.integers(0, h, 12), g.integers(0, w, 12)
    for y, x in zip(ys, xs):
        for b in range(B):
            ref = orc.esim_video_to_voxel(vids[b][:, y:y + 1, x:x + 1], pmap[b, y, x], nmap[b, y, x], 0.0,
                                          u0[b][y:y + 1, x:x + 1], np.zeros((1, 1)), np.zeros((n - 1, 1, 1)), False, lut)
            assert np.array_equal(got[b, :, :, y, x].reshape(-1).astype(np.float64), ref.reshape(-1))


def test_frame_out_padding_and_chunking(cuda_device):
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    n, h, w = 21, 36, 40                      # 36 is not a multiple of 16 -> padded rows
    vid = synth_video("walk", n, h, w, 9)
    u0 = np.random.Generator(np.random.PCG64(2)).random((h, w))
    fr = torch.from_numpy(vid).to(cuda_device)
    base = v2v.frames_to_voxel(fr, 0.2, 0.3, num_bins=5, frames_per_bin=2, u0=u0[None], lut=lut, return_potential=True)
    for mode, add in (("frames", False), ("frames+first", True)):
        o = v2v.frames_to_voxel(fr, 0.2, 0.3, num_bins=5, frames_per_bin=2, u0=u0[None], lut=lut, frame_out=mode,
                                pad_multiple=16)
        assert o.padded.shape[-2:] == (48, 48)
        assert torch.equal(o.voxel, base.voxel)
        assert float(o.padded[..., h:, :].abs().sum()) == 0 and float(o.padded[..., :, w:].abs().sum()) == 0
        ref = orc.pack_frames(vid[..., None], 10, 2, add)
        assert np.array_equal(o.frames[0].cpu().numpy(), ref)
    # chunked clip: carrying the potential across calls equals one long call
    a = v2v.frames_to_voxel(fr[:11], 0.2, 0.3, num_bins=5, frames_per_bin=2, u0=u0[None], lut=lut, return_potential=True)
    b = v2v.frames_to_voxel(fr[10:], 0.2, 0.3, num_bins=5, frames_per_bin=2, potential_in=a.potential, lut=lut,
                            return_potential=True)
    assert torch.equal(torch.cat([a.voxel, b.voxel], dim=1), base.voxel)
    assert torch.equal(b.potential, base.potential)


def test_philox_statistics_and_determinism(cuda_device):
    import v2v_b200 as v2v
    n, h, w = 41, 128, 128
    vid = np.full((n, h, w), 128, dtype=np.uint8)            # static scene: every event is noise
    fr = torch.from_numpy(vid).to(cuda_device)
    kw = dict(num_bins=1, noise="philox", base_noise_std=0.05, hot_pixel_fraction=0.01, hot_pixel_std=5.0,
              return_potential=True, with_stats=True)
    a = v2v.frames_to_voxel(fr, 0.2, 0.2, seed=123, **kw)
    b = v2v.frames_to_voxel(fr, 0.2, 0.2, seed=123, **kw)
    c = v2v.frames_to_voxel(fr, 0.2, 0.2, seed=124, **kw)
    assert torch.equal(a.voxel, b.voxel) and not torch.equal(a.voxel, c.voxel)
    # potential stays inside (-neg, pos) and initial potential is uniform: compare event rate with the oracle
    g = np.random.Generator(np.random.PCG64(0))
    u0, m = g.random((h, w)), g.random((h, w)) < 0.01
    hot = np.where(m, 5.0 * g.standard_normal((h, w)), 0.0)
    ref = orc.esim_video_to_voxel(vid, 0.2, 0.2, 0.05, u0, hot, g.standard_normal((n - 1, h, w)), False)
    rate_ref = np.abs(ref).sum() / ref.size
    rate = float(a.voxel.abs().sum()) / a.voxel.numel()
    assert abs(rate - rate_ref) / rate_ref < 0.15
    pot = a.potential.cpu().numpy()
    assert pot.max() < 0.2 and pot.min() > -0.2
    st = a.stats.cpu().numpy()[0]
    assert st[0] == int(a.voxel.clamp(min=0).sum()) and st[1] == int((-a.voxel).clamp(min=0).sum())
    # clip streams are distinct: same seed, shifted clip index
    d = v2v.frames_to_voxel(fr, 0.2, 0.2, seed=123, clip_index_base=1, **kw)
    assert not torch.equal(a.voxel, d.voxel)


def test_full_size_properties(cuda_device):
    """BASELINE config 2 size (121x480x640): size-independent properties instead of a full oracle run."""
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    n, h, w = 121, 480, 640
    vid = synth_video("walk", n, h, w, 3)
    fr = torch.from_numpy(vid).to(cuda_device)
    pos, neg = 0.23, 0.31
    g = np.random.Generator(np.random.PCG64(4))
    u0 = g.random((h, w))
    o = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, u0=u0[None], lut=lut, return_potential=True, with_stats=True)
    vox = o.voxel[0].to(torch.float64)
    # conservation: potential_end + pos*P - neg*N == potential_0 + L_end - L_0 (up to accumulated rounding)
    pe = vox.clamp(min=0).sum(dim=(0, 1)).cpu().numpy()
    ne = (-vox).clamp(min=0).sum(dim=(0, 1)).cpu().numpy()
    pot0 = u0 * (pos + neg) - neg
    lhs = o.potential[0].cpu().numpy() + pe * pos - ne * neg
    # intervals with both polarities summed inside one bin cannot occur with fpb=1, so pe/ne are exact
    assert np.allclose(lhs, pot0 + lut[vid[-1]] - lut[vid[0]], atol=1e-9)
    assert float(o.potential.max()) < pos and float(o.potential.min()) > -neg
    st = o.stats.cpu().numpy()[0]
    assert st[0] == int(pe.sum()) and st[1] == int(ne.sum())
    # a strip of the full-size result against the oracle
    sl = slice(200, 204)
    ref = orc.bin_accumulate(orc.esim_video_to_voxel(vid[:, sl], pos, neg, 0.0, u0[sl], np.zeros((4, w)),
                                                     np.zeros((n - 1, 4, w)), False, lut), 5, 1)
    assert np.array_equal(o.voxel[0, :, :, sl].cpu().numpy().astype(np.float64), ref)


def test_errors(cuda_device):
    import v2v_b200 as v2v
    fr = torch.zeros((1, 7, 8, 8), dtype=torch.uint8, device=cuda_device)
    with pytest.raises(AssertionError):
        v2v.frames_to_voxel(fr, 0.2, 0.2, num_bins=5)             # (N-1) % 5 != 0, data/v2v_datasets.py:365
    with pytest.raises(v2v.V2VError):
        v2v.frames_to_voxel(fr.cpu(), 0.2, 0.2, num_bins=1)       # no CPU fallback
    out = v2v.frames_to_voxel(torch.zeros((1, 1, 8, 8), dtype=torch.uint8, device=cuda_device), 0.2, 0.2, num_bins=5)
    assert out.voxel.shape == (1, 0, 5, 8, 8)                     # single frame: no intervals
    assert v2v.EventEmulator(rng="numpy").video_to_voxel(np.zeros((1, 4, 4), np.uint8)).shape == (0, 4, 4)


# ---------------------------------------------------------------------------------------------
# the throughput kernel (csrc/esim_fast.cu): taken for per-clip thresholds, fpb=1, noise none/philox,
# and B*H*W >= 148*2048 pixels with H*W % 4 == 0
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind,pos,neg", [("walk", 0.2, 0.2), ("iid", 0.05, 0.0625), ("iid", 1.3, 0.9), ("walk", 0.11, 0.1)])
def test_fast_kernel_noise_free_vs_oracle(cuda_device, kind, pos, neg):
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    n, h, w = 16, 480, 640                      # ragged trip count: 15 intervals = 3 trips of 4 + tail of 3
    vid = synth_video(kind, n, h, w, 21)
    g = np.random.Generator(np.random.PCG64(8))
    u0 = g.random((h, w))
    fr = torch.from_numpy(vid).to(cuda_device)
    o = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, u0=u0[None], lut=lut, with_stats=True, return_potential=True,
                            frame_out="frames")
    ref, pot = orc.esim_video_to_voxel(vid, pos, neg, 0.0, u0, np.zeros((h, w)), np.zeros((n - 1, h, w)), False, lut,
                                       return_state=True)
    assert np.array_equal(o.voxel[0].cpu().numpy().astype(np.float64), orc.bin_accumulate(ref, 5, 1))
    assert np.array_equal(o.potential[0].cpu().numpy(), pot)
    st = o.stats[0].cpu().numpy()
    assert st[0] == int(np.maximum(ref, 0).sum()) and st[1] == int(np.maximum(-ref, 0).sum())
    assert np.array_equal(o.frames[0].cpu().numpy(), orc.pack_frames(vid[..., None], 5, 3, False))
    # the generic kernel must agree bit for bit
    o2 = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, u0=u0[None], lut=lut, with_stats=True, return_potential=True,
                             kernel_flags=_lib.ESIM_FLAG_GENERIC)
    assert torch.equal(o.voxel, o2.voxel) and torch.equal(o.potential, o2.potential) and torch.equal(o.stats, o2.stats)


def test_philox_run_equals_oracle_on_dumped_fields(cuda_device):
    """Production mode end to end: the Philox kernels (fast and generic) == the CPU oracle fed with the very
    random fields the generator produced (dumped through the audit hook)."""
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    n, h, w = 11, 480, 640
    vid = synth_video("walk", n, h, w, 31)
    fr = torch.from_numpy(vid).to(cuda_device)
    pos, neg, std, frac, hstd = 0.18, 0.26, 0.07, 0.002, 6.0
    kw = dict(num_bins=5, noise="philox", base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=77,
              clip_index_base=5, with_stats=True, return_potential=True)
    fast = v2v.frames_to_voxel(fr, pos, neg, **kw)
    gen = v2v.frames_to_voxel(fr, pos, neg, kernel_flags=_lib.ESIM_FLAG_GENERIC, **kw)
    assert torch.equal(fast.voxel, gen.voxel) and torch.equal(fast.potential, gen.potential)
    assert torch.equal(fast.stats, gen.stats)
    u0, hot, bn = v2v.philox_fields(n, h, w, base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=77,
                                    clip_index_base=5)
    u0, hot, bn = u0[0].cpu().numpy(), hot[0].cpu().numpy(), bn[0].cpu().numpy()
    # sanity of the generator itself
    assert 0.0 <= u0.min() and u0.max() < 1.0 and abs(u0.mean() - 0.5) < 0.01
    nz = (hot != 0).mean()
    assert abs(nz - frac) < 0.0008
    z = bn / np.float32(std)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs((np.abs(z) > 3).mean() - 0.0027) < 0.0005
    ref, pot = orc.esim_video_to_voxel(vid, pos, neg, 1.0, u0, hot, bn, False, lut, return_state=True)
    assert np.array_equal(fast.voxel[0].cpu().numpy().astype(np.float64), orc.bin_accumulate(ref, 5, 1))
    assert np.array_equal(fast.potential[0].cpu().numpy(), pot)
    # and explicit replay on the GPU
    rep = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, noise="explicit", base_noise_std=1.0, u0=u0[None], hot_noise=hot[None],
                              base_gauss=bn[None], lut=lut)
    assert torch.equal(rep.voxel, fast.voxel)


def test_host_pipeline_and_batch_api(cuda_device):
    """Public host-buffer API == direct batched call; batch_to_tensors returns the train-loop layout."""
    import v2v_b200 as v2v
    B, n, h, w = 5, 11, 480, 640
    vids = np.stack([synth_video("walk", n, h, w, 70 + b) for b in range(B)])
    vz = v2v.V2VVoxelizer(dict(num_bins=5, base_noise_std_range=[0, 0.1], hot_pixel_std_range=[0, 10]), device=cuda_device)
    params = vz.sample_batch_params(B, rs=np.random.RandomState(3))
    fr = torch.from_numpy(vids).to(cuda_device)
    ref = vz.batch_to_tensors(fr, params, seed=9, clip_index_base=40, with_stats=True)
    assert ref["events"].shape == (B, 2, 5, h, w) and ref["frame"].shape == (B, 2, 1, h, w)
    assert ref["events"].dtype == torch.float32 and ref["frame"].dtype == torch.float32
    assert np.array_equal(ref["frame"][1].cpu().numpy(), orc.pack_frames(vids[1][..., None], 5, 2, False))
    # padded consumer layout: same values, zero pads
    pad = vz.batch_to_tensors(fr, params, seed=9, clip_index_base=40, pad_multiple=16 * 31)     # 480 -> 496, 640 -> 992
    assert pad["events_padded"].shape[-2:] == (496, 992) and torch.equal(pad["events"], ref["events"])
    assert float(pad["events_padded"][..., 480:, :].abs().sum()) == 0.0
    # host pipeline (pinned in / pinned out, 2 clips per chunk -> ragged last chunk)
    host_in = torch.from_numpy(vids).pin_memory()
    host_out = torch.empty((B, 2, 5, h, w), dtype=torch.float32).pin_memory()
    pipe = v2v.HostPipeline(vz, cuda_device, clips_per_chunk=2, seed=9)
    st = pipe.run(host_in, params, host_out, clip_index_base=40)
    torch.cuda.synchronize()
    assert torch.equal(host_out, ref["events"].cpu())
    assert torch.equal(st.cpu(), ref["stats"].cpu())
    # back-to-back calls overlap (the second call's copies and kernels run under the first call's draining D2H stream) and
    # reuse the staging slots: both results must still be whole after synchronize()
    host_b = torch.empty_like(host_out).pin_memory()
    host_out.zero_()
    ref_b = vz.batch_to_tensors(fr, params, seed=9, clip_index_base=900, with_stats=True)
    pipe.run(host_in, params, host_out, clip_index_base=40)
    st_b = pipe.run(host_in, params, host_b, clip_index_base=900)
    pipe.synchronize()
    assert torch.equal(host_out, ref["events"].cpu()) and torch.equal(host_b, ref_b["events"].cpu())
    assert torch.equal(st_b.cpu(), ref_b["stats"].cpu())
    # sharding invariance: clips simulated one by one with their global index give the same noise streams
    one = vz.batch_to_tensors(fr[3:4], params[3:4], seed=9, clip_index_base=43)
    assert torch.equal(one["events"][0], ref["events"][3])


def test_full_size_clip_bit_exact_vs_c_oracle(cuda_device):
    """BASELINE config 2 clip (121x480x640) in the production mode (Philox noise, fast kernel): every one of the
    36.9 M pixel-intervals equals the C oracle fed with the dumped noise fields; potential and stats too."""
    import v2v_b200 as v2v
    import v2v_oracle_c as orcc
    lut = orc.esim_log_lut()
    n, h, w = 121, 480, 640
    vid = synth_video("walk", n, h, w, 1234)
    fr = torch.from_numpy(vid).to(cuda_device)
    pos, neg, std, frac, hstd = 0.21, 0.29, 0.06, 0.0007, 8.0
    o = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, noise="philox", base_noise_std=std, hot_pixel_fraction=frac,
                            hot_pixel_std=hstd, seed=5, clip_index_base=17, with_stats=True, return_potential=True)
    u0, hot, bn = v2v.philox_fields(n, h, w, base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=5,
                                    clip_index_base=17)
    ref, pot = orcc.esim_video_to_voxel(vid, pos, neg, 1.0, u0[0].cpu().numpy(), hot[0].cpu().numpy(), bn[0].cpu().numpy(),
                                        False, lut, return_state=True)
    got = o.voxel[0].cpu().numpy().reshape(n - 1, h, w)
    assert bool(torch.isfinite(bn).all())           # (an SFU lg2 a hair above -2 once gave NaN radii)
    assert np.array_equal(got.astype(np.float64), ref)
    assert np.array_equal(o.potential[0].cpu().numpy(), pot)
    st = o.stats[0].cpu().numpy()
    assert st[0] == int(np.maximum(ref, 0).sum()) and st[1] == int(np.maximum(-ref, 0).sum())


def test_config5_shape_bit_exact_vs_c_oracle(cuda_device):
    """BASELINE config 5 shape: a 1080p clip of 26 frames (5 voxels of 5 bins) in the production mode (Philox noise,
    statistics, ground-truth frames, voxels written into the consumer's /16-padded layout, model/train_utils.py:322-326):
    all 51.8 M pixel-intervals equal the C oracle fed with the dumped noise fields; pads stay zero."""
    import v2v_b200 as v2v
    import v2v_oracle_c as orcc
    lut = orc.esim_log_lut()
    n, h, w = 26, 1080, 1920
    vid = synth_video("walk", n, h, w, 77)
    fr = torch.from_numpy(vid).to(cuda_device)
    pos, neg, std, frac, hstd = 0.33, 0.27, 0.04, 0.0005, 4.0
    o = v2v.frames_to_voxel(fr, pos, neg, num_bins=5, noise="philox", base_noise_std=std, hot_pixel_fraction=frac,
                            hot_pixel_std=hstd, seed=9, clip_index_base=3, with_stats=True, return_potential=True,
                            pad_multiple=16, frame_out="frames")
    assert o.padded.shape == (1, 5, 5, 1088, 1920) and o.voxel.shape == (1, 5, 5, h, w)
    assert float(o.padded[..., h:, :].abs().sum()) == 0.0
    u0, hot, bn = v2v.philox_fields(n, h, w, base_noise_std=std, hot_pixel_fraction=frac, hot_pixel_std=hstd, seed=9,
                                    clip_index_base=3)
    assert bool(torch.isfinite(bn).all())
    ref, pot = orcc.esim_video_to_voxel(vid, pos, neg, 1.0, u0[0].cpu().numpy(), hot[0].cpu().numpy(), bn[0].cpu().numpy(),
                                        False, lut, return_state=True)
    got = o.voxel[0].cpu().numpy().reshape(n - 1, h, w)
    assert np.array_equal(got.astype(np.float64), ref)
    assert np.array_equal(o.potential[0].cpu().numpy(), pot)
    st = o.stats[0].cpu().numpy()
    assert st[0] == int(np.maximum(ref, 0).sum()) and st[1] == int(np.maximum(-ref, 0).sum())
    assert np.array_equal(o.frames[0].cpu().numpy(), orc.pack_frames(vid[..., None], 5, 5, False))


def test_train_batch_shape_bit_exact_vs_c_oracle(cuda_device):
    """The batch the shipped training config builds (config/train_v2v_e2vid_10k.yaml:62-71: 12 clips of 201 frames,
    128x128 crops, 40 voxels of 5 bins): Philox mode with per-clip parameters drawn by the reference's law, every clip
    against the C oracle on its dumped fields; the small-launch kernel choice must not change a bit."""
    import v2v_b200 as v2v
    import v2v_oracle_c as orcc
    lut = orc.esim_log_lut()
    B, n, h, w = 12, 201, 128, 128
    vids = np.stack([synth_video("walk", n, h, w, 300 + b) for b in range(B)])
    vz = v2v.V2VVoxelizer(dict(num_bins=5, base_noise_std_range=[0, 0.1], hot_pixel_std_range=[0, 10]), device=cuda_device)
    params = vz.sample_batch_params(B, rs=np.random.RandomState(8))
    col = lambda k: np.array([p[k] for p in params])
    fr = torch.from_numpy(vids).to(cuda_device)
    kw = dict(num_bins=5, noise="philox", base_noise_std=col("base_noise_std"), hot_pixel_fraction=col("hot_pixel_fraction"),
              hot_pixel_std=col("hot_pixel_std"), seed=31, clip_index_base=100, with_stats=True, return_potential=True)
    o = v2v.frames_to_voxel(fr, col("pos_thres"), col("neg_thres"), **kw)
    for flags in (_lib.ESIM_FLAG_GENERIC, _lib.ESIM_FLAG_SMALL_P1, _lib.ESIM_FLAG_SMALL_FAST):
        o2 = v2v.frames_to_voxel(fr, col("pos_thres"), col("neg_thres"), kernel_flags=flags, **kw)
        assert torch.equal(o.voxel, o2.voxel) and torch.equal(o.potential, o2.potential) and torch.equal(o.stats, o2.stats)
    u0, hot, bn = v2v.philox_fields(n, h, w, base_noise_std=col("base_noise_std"), hot_pixel_fraction=col("hot_pixel_fraction"),
                                    hot_pixel_std=col("hot_pixel_std"), seed=31, clip_index_base=100)
    for b in range(B):
        ref, pot = orcc.esim_video_to_voxel(vids[b], params[b]["pos_thres"], params[b]["neg_thres"], 1.0, u0[b].cpu().numpy(),
                                            hot[b].cpu().numpy(), bn[b].cpu().numpy(), False, lut, return_state=True)
        got = o.voxel[b].cpu().numpy().reshape(n - 1, h, w)
        assert np.array_equal(got.astype(np.float64), ref), f"clip {b}"
        assert np.array_equal(o.potential[b].cpu().numpy(), pot)
        st = o.stats[b].cpu().numpy()
        assert st[0] == int(np.maximum(ref, 0).sum()) and st[1] == int(np.maximum(-ref, 0).sum())


def test_staged_frame_ring_variant_is_bit_identical(cuda_device):
    """The noise-free throughput kernel with its frames staged through shared memory by bulk async copies
    (cp.async.bulk + mbarrier ring, V2V_ESIM_FLAG_STAGED) equals the per-lane LDG form bit for bit: full tiles, a ragged
    last tile, statistics and ground-truth frames, interval counts that are not multiples of the ring depth."""
    import v2v_b200 as v2v
    lut = orc.esim_log_lut()
    for (B, n, h, w) in ((3, 21, 480, 640), (2, 11, 72, 112), (1, 6, 480, 644)):
        vids = np.stack([synth_video("walk", n, h, w, 500 + b) for b in range(B)])
        fr = torch.from_numpy(vids).to(cuda_device)
        u0 = torch.rand((B, h, w), dtype=torch.float64, device=cuda_device)
        kw = dict(num_bins=5, u0=u0, lut=lut, with_stats=True, return_potential=True, frame_out="frames",
                  kernel_flags=_lib.ESIM_FLAG_SMALL_FAST)
        a = v2v.frames_to_voxel(fr, [0.21] * B, [0.17] * B, **kw)
        kw["kernel_flags"] |= _lib.ESIM_FLAG_STAGED
        b_ = v2v.frames_to_voxel(fr, [0.21] * B, [0.17] * B, **kw)
        assert torch.equal(a.voxel, b_.voxel) and torch.equal(a.potential, b_.potential)
        assert torch.equal(a.stats, b_.stats) and torch.equal(a.frames, b_.frames)
        ref = orc.esim_video_to_voxel(vids[0], 0.21, 0.17, 0.0, u0[0].cpu().numpy(), np.zeros((h, w)), np.zeros((n - 1, h, w)), False, lut)
        assert np.array_equal(b_.voxel[0].cpu().numpy().reshape(n - 1, h, w).astype(np.float64), ref)


def test_noise_field_distribution_and_independence(cuda_device):
    """Statistical audit of the in-kernel generator (Philox-seeded 64-bit LCG streams + table Box-Muller) on the dumped
    base-noise field: moments, Kolmogorov distance to the normal law, tails, and the correlations the construction could
    plausibly introduce (the two normals of a Box-Muller pair = neighbouring pixels, the two pairs of a group =
    consecutive stream words, consecutive intervals = consecutive draws of one stream, neighbouring groups = different
    streams with different direction sub-tables, different clips)."""
    from scipy import stats as sst
    from v2v_b200.esim import philox_fields
    n, h, w = 65, 96, 128
    _, _, bn = philox_fields(n, h, w, base_noise_std=[1.0, 1.0], hot_pixel_fraction=[0.0, 0.0], hot_pixel_std=[0.0, 0.0], seed=77)
    z = bn.cpu().numpy()                                   # [2, 64, h, w], unit variance
    x = z.ravel()
    m = x.size
    assert abs(x.mean()) < 5 / np.sqrt(m)
    assert abs(x.var() - 1) < 5 * np.sqrt(2 / m)
    assert abs(sst.skew(x)) < 5 * np.sqrt(6 / m)
    assert abs(sst.kurtosis(x)) < 5 * np.sqrt(24 / m) + 2e-3      # 21-bit radius: the tail is cut at 5.40 sigma
    sub = x[:: 7][:400_000]
    assert sst.kstest(sub, "norm").statistic < 1.95 / np.sqrt(sub.size)   # alpha ~ 1e-3
    for t in (1.0, 2.0, 3.0, 4.0):                          # two-sided tail mass
        p = 2 * sst.norm.sf(t)
        assert abs((np.abs(x) > t).mean() - p) < 5 * np.sqrt(p / m) + 1e-7
    assert np.abs(x).max() < 5.41
    # every group uses its own 256 of the 2048 directions: the marginal law of each of the 8 classes is normal too
    grp = (np.arange(h * w) // 4) & 7
    zc = z.reshape(2, n - 1, h * w)
    for c in range(8):
        xc = zc[..., grp == c].ravel()
        assert abs(xc.var() - 1) < 5 * np.sqrt(2 / xc.size) and abs(sst.kurtosis(xc)) < 5 * np.sqrt(24 / xc.size) + 2e-3
        assert sst.kstest(xc[:: 3][:300_000], "norm").statistic < 1.95 / np.sqrt(min(300_000, xc[:: 3].size))

    def corr(a, b):
        return float(np.corrcoef(a.ravel(), b.ravel())[0, 1])

    lim = 5 / np.sqrt(m / 2)
    assert abs(corr(z[..., 0::2], z[..., 1::2])) < lim                               # the two normals of one Box-Muller pair
    assert abs(corr(z[..., 0::2] ** 2, z[..., 1::2] ** 2)) < lim                      # ... and their magnitudes
    assert abs(corr(z[..., 1::4], z[..., 2::4])) < lim                               # the two words of one interval
    assert abs(corr(z[:, :-1], z[:, 1:])) < lim and abs(corr(z[:, :-1] ** 2, z[:, 1:] ** 2)) < lim   # consecutive intervals
    assert abs(corr(z[:, :-2], z[:, 2:])) < lim
    assert abs(corr(z[..., 3:-4:4], z[..., 4::4])) < lim                             # neighbouring groups (different streams)
    assert abs(corr(z[..., :-1, :], z[..., 1:, :])) < lim                            # neighbouring rows
    assert abs(corr(z[0], z[1])) < lim                                                # different clips, same seed


def test_rng_known_answers_and_noise_chain(cuda_device):
    """The device generators are the documented algorithms (Philox4x32-10 KATs of Random123; the base-noise stream ==
    the Philox-seeded 64-bit LCG of the oracle's restatement, word for word; the direction table == the oracle's), and
    the dumped noise field is the documented function of those words up to the SFU approximations of the radius."""
    import ctypes as C
    from v2v_b200 import _lib
    from v2v_b200.esim import philox_fields
    lib = _lib.load()
    s = torch.cuda.current_stream(cuda_device).cuda_stream

    def rng_words(counter, key, seed, clip, group, n):
        po = torch.zeros(4, dtype=torch.int32, device=cuda_device)
        wo = torch.zeros(max(n, 1), dtype=torch.int32, device=cuda_device)
        _lib.check(lib.v2v_rng_words((C.c_uint32 * 4)(*counter), (C.c_uint32 * 2)(*key), C.c_void_p(po.data_ptr()), seed, clip, group, n,
                                     C.c_void_p(wo.data_ptr()), C.c_void_p(s)))
        u = lambda t: [int(x) & 0xFFFFFFFF for x in t.cpu().tolist()]
        return u(po), u(wo)[:n]

    for ctr, key in (([0] * 4, [0] * 2), ([0xffffffff] * 4, [0xffffffff] * 2),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])):
        assert rng_words(ctr, key, 0, 0, 0, 0)[0] == orc.philox4x32_10(ctr, key)
    for seed, clip, group in ((0, 0, 0), (123, 5, 77), (2 ** 63 + 12345, 2 ** 40 + 3, 2 ** 33 + 9), (77, 1, 3071)):
        assert rng_words([0] * 4, [0] * 2, seed, clip, group, 64)[1] == orc.esim_noise_stream_words(seed, clip, group, 64)
    # the field: clip 1 of a 2-clip dump, pixel groups of a 96x128 frame, first 6 intervals
    h, w, n, std, seed = 96, 128, 7, 0.37, 77
    _, _, bn = philox_fields(n, h, w, base_noise_std=[std, std], hot_pixel_fraction=[0.0, 0.0], hot_pixel_std=[0.0, 0.0], seed=seed)
    z = bn[1].cpu().numpy().reshape(n - 1, h * w)
    table = orc.esim_direction_table()
    for group in (0, 1, 5, 500, 3071):
        words = orc.esim_noise_stream_words(seed, 1, group, 2 * (n - 1))
        want = orc.esim_noise_from_words(words, std, group, table)                    # [n-1, 4]
        got = z[:, 4 * group: 4 * group + 4]
        assert np.allclose(got, want, rtol=2e-4, atol=2e-6 * std)
        # every value is an exact product of a float32 radius and a 21-bit direction of this group's sub-table
        idx = ((np.asarray(words, dtype=np.int64) >> 21) & 0x7F8) | (group & 7)
        r = got.reshape(-1, 2) / table[idx]
        assert np.array_equal(r[:, 0].astype(np.float32).astype(np.float64), r[:, 0]) and np.allclose(r[:, 0], r[:, 1], rtol=1e-15)
    # a zero std gives an exactly zero field
    _, _, b0 = philox_fields(3, 8, 16, base_noise_std=[0.0], hot_pixel_fraction=[0.0], hot_pixel_std=[0.0], seed=1)
    assert float(b0.abs().max()) == 0.0
