"""GPU parity of the v2e-style kernel against the reference's golden outputs (counts bit-exact)."""
import numpy as np
import pytest
import torch

import v2v_oracle as orc
from conftest import golden, synth_video

pytestmark = pytest.mark.gpu


def params_of(c):
    p = {k[2:]: float(v) for k, v in c.items() if k.startswith("p_")}
    p["threshold_model"] = str(c["threshold_model"])
    return p


@pytest.mark.parametrize("name", golden("v2e").cases)
def test_v2e_golden_replay(cuda_device, name):
    """Explicit random fields recorded from the reference run -> identical counts."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    c = golden("v2e").case(name)
    p = params_of(c)
    fr = torch.from_numpy(c["video"]).to(cuda_device)
    per_frame = "pos_thres_frames" in c        # spatial_temporal_independent: the maps in force at every frame
    out = frames_to_voxel_v2e(
        fr, c["pos_thres_frames" if per_frame else "pos_thres"][None], c["neg_thres_frames" if per_frame else "neg_thres"][None],
        fps=float(c["fps"]), cutoff_hz=p["cutoff_hz"],
        leak_rate_hz=p["leak_rate_hz"], shot_noise_rate_hz=p["shot_noise_rate_hz"],
        leak_jitter_fraction=p["leak_jitter_fraction"], noise_rate=c["noise_rate"][None],
        pos_thres_nominal=p["thres_mean_mean"] + p["thres_diff_mean"] / 2,
        neg_thres_nominal=p["thres_mean_mean"] - p["thres_diff_mean"] / 2, noise="explicit",
        leak_randn=c["leak_randn"][None] if "leak_randn" in c else None,
        pos_shot=c["pos_shot"][None] if "pos_shot" in c else None,
        neg_shot=c["neg_shot"][None] if "neg_shot" in c else None, lut=c["lut"], with_stats=True,
        u8_intensity=bool(int(c.get("u8_input", 0))))
    got = out["voxel"][0, :, 0].cpu().numpy().astype(np.float64)
    assert np.array_equal(got, c["ref"])


@pytest.mark.parametrize("name", golden("v2e").cases)
def test_v2e_reference_signature_same_seed(cuda_device, name):
    """video_to_voxel(...) with rng='numpy' and the reference's seed reproduces the reference."""
    from v2v_b200.v2e import video_to_voxel
    c = golden("v2e").case(name)
    p = params_of(c)
    video = c["video"] if int(c.get("u8_input", 0)) else c["video"].astype(np.float64)
    got = video_to_voxel(video, int(c["fps"]), refractory_period_s=0, seed=int(c["seed"]),
                         rng="numpy", lut=c["lut"], **p)
    assert got.dtype == np.float64 and np.array_equal(got, c["ref"])


def test_v2e_large_vectorised_vs_oracle(cuda_device):
    """Config-3 style: HDR-degraded clip at a size that takes the 4-pixel path; noisy preset replayed."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    n, h, w = 11, 480, 640
    vid = synth_video("walk", n, h, w, 77)
    vid = np.clip((vid - 127.5) * 2.3 + 127.5, 0, 255).astype(np.uint8)         # data/v2v_datasets.py:473-477
    p = dict(threshold_model="pn_related", thres_mean_mean=0.2, thres_mean_std=0.05, thres_diff_mean=0.0,
             thres_diff_std=0.05, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1,
             noise_rate_cov_decades=0.1)
    rec = {}
    np.random.seed(5)
    ref = orc.v2e_video_to_voxel(vid.astype(np.float64), 24, p, np.random, record=rec)
    fr = torch.from_numpy(vid).to(cuda_device)
    out = frames_to_voxel_v2e(fr, rec["pos_thres"][None], rec["neg_thres"][None], fps=24, cutoff_hz=30.0, leak_rate_hz=0.1,
                              shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1, noise_rate=rec["noise_rate"][None],
                              pos_thres_nominal=0.2, neg_thres_nominal=0.2, noise="explicit",
                              leak_randn=np.stack(rec["leak_randn"])[None],
                              pos_shot=np.stack(rec["pos_shot"]).astype(np.int32)[None],
                              neg_shot=np.stack(rec["neg_shot"]).astype(np.int32)[None])
    assert np.array_equal(out["voxel"][0, :, 0].cpu().numpy().astype(np.float64), ref)


def test_v2e_philox_statistics(cuda_device):
    """In-kernel Philox leak jitter + Poisson shot noise: deterministic per seed; the run equals the CPU oracle
    replaying the very fields the generator drew (audit hook); the fields have the right distribution."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    n, h, w = 13, 256, 256
    vid = synth_video("walk", n, h, w, 8)
    fr = torch.from_numpy(vid).to(cuda_device)
    p = dict(threshold_model="pn_related", thres_mean_mean=0.2, thres_mean_std=0.03, thres_diff_mean=0.0,
             thres_diff_std=0.03, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1,
             noise_rate_cov_decades=0.1)
    g = np.random.Generator(np.random.PCG64(1))
    pos = np.clip(g.normal(0.2, 0.03, (h, w)), 0.01, None)
    neg = np.clip(g.normal(0.2, 0.03, (h, w)), 0.01, None)
    nrate = np.exp(np.log(10) * 0.1 * g.standard_normal((h, w)).astype(np.float32)).astype(np.float32)
    kw = dict(fps=24, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1,
              noise_rate=nrate[None], noise="philox", with_stats=True)
    a = frames_to_voxel_v2e(fr, pos[None], neg[None], seed=1, return_fields=True, **kw)
    b = frames_to_voxel_v2e(fr, pos[None], neg[None], seed=1, **kw)
    c = frames_to_voxel_v2e(fr, pos[None], neg[None], seed=2, **kw)
    assert torch.equal(a["voxel"], b["voxel"]) and not torch.equal(a["voxel"], c["voxel"])
    f = {k: v[0].cpu().numpy() for k, v in a["fields"].items()}
    # distribution of the drawn fields
    z = f["leak_randn"]
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    lam = 2.5 / 24                                      # rate/2 * dt, averaged over the frame by construction
    assert abs(f["pos_shot"].mean() - lam) / lam < 0.03 and abs(f["neg_shot"].mean() - lam) / lam < 0.03
    # oracle replay of exactly these fields (maps given, per-frame draws replayed in the reference's order)
    class Replay:
        def __init__(self):
            self.leak, self.shots = list(f["leak_randn"]), []
            for x, y in zip(f["pos_shot"], f["neg_shot"]):
                self.shots += [x, y]

        def randn(self, *shape):
            return self.leak.pop(0)

        def poisson(self, lam):
            return self.shots.pop(0)

    ref = orc.v2e_video_to_voxel(vid.astype(np.float64), 24, p, Replay(), maps=(pos, neg, nrate))
    assert np.array_equal(a["voxel"][0, :, 0].cpu().numpy().astype(np.float64), ref)
    # explicit replay on the GPU agrees too
    e = frames_to_voxel_v2e(fr, pos[None], neg[None], fps=24, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0,
                            leak_jitter_fraction=0.1, noise_rate=nrate[None], noise="explicit",
                            leak_randn=f["leak_randn"][None], pos_shot=f["pos_shot"][None], neg_shot=f["neg_shot"][None])
    assert torch.equal(e["voxel"], a["voxel"])


# ---- throughput kernel (csrc/v2e_fast.cu): forced on small shapes with V2V_V2E_FAST=1 ---------------------------

_FAST_NONE = {
    "clean_f32state": dict(cutoff_hz=0.0, leak_rate_hz=0.0),
    "cutoff": dict(cutoff_hz=30.0, leak_rate_hz=0.0),
    "leak": dict(cutoff_hz=0.0, leak_rate_hz=0.3),
    "cutoff_leak": dict(cutoff_hz=15.0, leak_rate_hz=0.2),
}


@pytest.mark.parametrize("bf", ["1", "0"])
@pytest.mark.parametrize("preset", sorted(_FAST_NONE))
@pytest.mark.parametrize("shape", [(13, 64, 96, 1, 1), (8, 40, 52, 1, 1), (13, 32, 36, 3, 2), (2, 16, 16, 1, 1)])
def test_v2e_fast_kernel_noise_free_vs_oracle(cuda_device, monkeypatch, preset, shape, bf):
    """Low-pass / leak / float32-state arithmetic of the throughput kernel == NumPy oracle, counts bit-exact
    (ragged interval counts, frames_per_bin > 1, hard HDR-degraded contrast so that multi-threshold crossings occur)."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    monkeypatch.setenv("V2V_V2E_FAST", "1")
    monkeypatch.setenv("V2V_V2E_BF", bf)          # exact division on every pixel / single-crossing fast path + divergent exact path
    n, h, w, bins, fpb = shape
    kw = _FAST_NONE[preset]
    vid = synth_video("walk", n, h, w, 31)
    vid = np.clip((vid - 127.5) * 2.6 + 127.5, 0, 255).astype(np.uint8)
    p = dict(threshold_model="pn_related", thres_mean_mean=0.2, thres_mean_std=0.05, thres_diff_mean=0.0, thres_diff_std=0.05,
             shot_noise_rate_hz=0.0, leak_jitter_fraction=0.0, noise_rate_cov_decades=0.1, **kw)
    rec = {}
    np.random.seed(11)
    ref = orc.v2e_video_to_voxel(vid.astype(np.float64), 24, p, np.random, record=rec)          # [n-1,h,w]
    fr = torch.from_numpy(vid).to(cuda_device)
    before = _launches()
    out = frames_to_voxel_v2e(fr, rec["pos_thres"][None], rec["neg_thres"][None], fps=24, num_bins=bins, frames_per_bin=fpb,
                              noise_rate=rec["noise_rate"][None], pos_thres_nominal=0.2, neg_thres_nominal=0.2, noise="none",
                              with_stats=True, **kw)
    assert _launches() == before + 1
    got = out["voxel"][0].cpu().numpy().astype(np.float64)                                       # [T,bins,h,w]
    T = (n - 1) // (bins * fpb)
    assert np.array_equal(got, ref.reshape(T, bins, fpb, h, w).sum(axis=2))
    # generic kernel on the same input: same voxels and the same event totals
    monkeypatch.delenv("V2V_V2E_FAST")
    monkeypatch.setenv("V2V_V2E_GENERIC", "1")
    gen = frames_to_voxel_v2e(fr, rec["pos_thres"][None], rec["neg_thres"][None], fps=24, num_bins=bins, frames_per_bin=fpb,
                              noise_rate=rec["noise_rate"][None], pos_thres_nominal=0.2, neg_thres_nominal=0.2, noise="none",
                              with_stats=True, **kw)
    assert torch.equal(gen["voxel"], out["voxel"]) and torch.equal(gen["stats"], out["stats"])
    if fpb == 1:
        assert int(out["stats"].sum()) == int(np.abs(ref).sum())


def _launches():
    from v2v_b200 import _lib
    return int(_lib.load().v2v_launch_count())


_FAST_PHILOX = {
    "noisy": dict(cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1),
    "leak_jitter": dict(cutoff_hz=0.0, leak_rate_hz=0.5, shot_noise_rate_hz=0.0, leak_jitter_fraction=0.3),
    "cutoff_shot": dict(cutoff_hz=20.0, leak_rate_hz=0.0, shot_noise_rate_hz=40.0, leak_jitter_fraction=0.0),
    "leak_shot_heavy": dict(cutoff_hz=0.0, leak_rate_hz=0.2, shot_noise_rate_hz=400.0, leak_jitter_fraction=0.1),
}


@pytest.mark.parametrize("bf", ["1", "0"])
@pytest.mark.parametrize("preset", sorted(_FAST_PHILOX))
@pytest.mark.parametrize("n", [13, 8])
def test_v2e_fast_kernel_philox_equals_generic_and_replay(cuda_device, monkeypatch, preset, n, bf):
    """Philox mode: throughput kernel == generic kernel == explicit replay of the dumped fields (bit for bit);
    `leak_shot_heavy` pushes the Poisson rate to ~8 events per frame so the k >= 3 tail loop runs everywhere."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    h, w = 48, 64
    kw = _FAST_PHILOX[preset]
    vid = synth_video("walk", n, h, w, 5)
    vid = np.clip((vid - 127.5) * 2.0 + 127.5, 0, 255).astype(np.uint8)
    fr = torch.from_numpy(np.stack([vid, vid[::-1].copy()])).to(cuda_device)                    # B = 2
    g = np.random.Generator(np.random.PCG64(2))
    pos = np.clip(g.normal(0.2, 0.05, (2, h, w)), 0.01, None)
    neg = np.clip(g.normal(0.2, 0.05, (2, h, w)), 0.01, None)
    nrate = np.exp(np.log(10) * 0.1 * g.standard_normal((2, h, w)).astype(np.float32)).astype(np.float32)
    common = dict(fps=24, noise_rate=nrate, with_stats=True, **kw)
    monkeypatch.setenv("V2V_V2E_FAST", "1")
    monkeypatch.setenv("V2V_V2E_BF", bf)
    fast = frames_to_voxel_v2e(fr, pos, neg, noise="philox", seed=9, clip_index_base=3, return_fields=True, **common)
    monkeypatch.delenv("V2V_V2E_FAST")
    monkeypatch.setenv("V2V_V2E_GENERIC", "1")
    gen = frames_to_voxel_v2e(fr, pos, neg, noise="philox", seed=9, clip_index_base=3, **common)
    assert torch.equal(fast["voxel"], gen["voxel"]) and torch.equal(fast["stats"], gen["stats"])
    f = fast["fields"]
    rep = frames_to_voxel_v2e(fr, pos, neg, noise="explicit",
                              leak_randn=f["leak_randn"] if kw["leak_rate_hz"] > 0 else None,
                              pos_shot=f["pos_shot"] if kw["shot_noise_rate_hz"] > 0 else None,
                              neg_shot=f["neg_shot"] if kw["shot_noise_rate_hz"] > 0 else None, **common)
    assert torch.equal(fast["voxel"], rep["voxel"]) and torch.equal(fast["stats"], rep["stats"])
    if kw["shot_noise_rate_hz"] > 0:
        lam = kw["shot_noise_rate_hz"] / 2 / 24
        m = float(f["pos_shot"].double().mean())
        assert abs(m - lam) / lam < 0.05
        if preset == "leak_shot_heavy":
            assert int(f["pos_shot"].max()) >= 12


def test_v2e_fast_kernel_full_size_philox_replay(cuda_device):
    """Config-3 shape (one HDR-degraded 480x640 clip, noisy preset): the throughput kernel is the default path;
    its run equals the explicit replay of the fields it drew through the generic kernel."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    n, h, w = 21, 480, 640
    vid = synth_video("walk", n, h, w, 77)
    vid = np.clip((vid - 127.5) * 2.3 + 127.5, 0, 255).astype(np.uint8)
    fr = torch.from_numpy(vid).to(cuda_device)
    g = np.random.Generator(np.random.PCG64(4))
    pos = np.clip(g.normal(0.2, 0.05, (1, h, w)), 0.01, None)
    neg = np.clip(g.normal(0.2, 0.05, (1, h, w)), 0.01, None)
    nrate = np.exp(np.log(10) * 0.1 * g.standard_normal((1, h, w)).astype(np.float32)).astype(np.float32)
    common = dict(fps=24, num_bins=5, cutoff_hz=30.0, leak_rate_hz=0.1, shot_noise_rate_hz=5.0, leak_jitter_fraction=0.1,
                  noise_rate=nrate, with_stats=True)
    a = frames_to_voxel_v2e(fr, pos, neg, noise="philox", seed=21, return_fields=True, **common)
    f = a["fields"]
    e = frames_to_voxel_v2e(fr, pos, neg, noise="explicit", leak_randn=f["leak_randn"], pos_shot=f["pos_shot"],
                            neg_shot=f["neg_shot"], **common)
    assert torch.equal(a["voxel"], e["voxel"]) and torch.equal(a["stats"], e["stats"])
    assert int(a["stats"].sum()) > 0


@pytest.mark.parametrize("shape", [(2, 9, 48, 64), (1, 5, 30, 34), (3, 4, 480, 640)])
def test_v2e_shot_scales_match_numpy(cuda_device, shape):
    """Per-frame Poisson normalisers (data/v2v_core_v2e.py:90-99): (rate/2*dt) / mean(inten_factor * nominal/thres),
    fixed-point accumulation -> equal to the float64 NumPy means to ~1e-10 and bit-identical from run to run."""
    from v2v_b200.v2e import frames_to_voxel_v2e
    B, n, h, w = shape
    g = np.random.Generator(np.random.PCG64(12))
    vid = g.integers(0, 256, (B, n, h, w), dtype=np.uint8)
    pos = np.clip(g.normal(0.2, 0.05, (B, h, w)), 0.01, None)
    neg = np.clip(g.normal(0.25, 0.05, (B, h, w)), 0.01, None)
    fr = torch.from_numpy(vid).to(cuda_device)
    kw = dict(fps=24, cutoff_hz=10.0, shot_noise_rate_hz=5.0, pos_thres_nominal=0.2, neg_thres_nominal=0.25, noise="philox", seed=1)
    a = frames_to_voxel_v2e(fr, pos, neg, **kw)
    b = frames_to_voxel_v2e(fr, pos, neg, **kw)
    assert torch.equal(a["shot_scales"], b["shot_scales"]) and torch.equal(a["voxel"], b["voxel"])
    fac = 1 - 0.75 * ((vid[:, 1:].astype(np.float64) + 20) / 275.0)                      # [B,n-1,h,w]
    k = np.arange(1, n, dtype=np.float64)
    dt = k / 24 - (k - 1) / 24
    for pol, thr, nom in ((0, pos, 0.2), (1, neg, 0.25)):
        mean = (fac * (nom / thr)[:, None]).mean(axis=(2, 3))
        ref = (5.0 / 2) * dt[None] / mean
        got = a["shot_scales"][pol].cpu().numpy()
        assert np.allclose(got, ref, rtol=1e-9, atol=0)
