"""GPU parity of the fused frame-side packing (pause gather + degrade inside the ESIM kernels) and of bgr_to_gray."""
import numpy as np
import pytest
import torch

import v2v_oracle as orc
from conftest import golden, synth_video

pytestmark = pytest.mark.gpu


def test_bgr_to_gray_kernel_matches_reference(cuda_device):
    import v2v_b200 as v2v
    c = golden("frames").case("bgr_to_gray")
    got = v2v.bgr_to_gray(torch.from_numpy(c["img"]).to(cuda_device)).cpu().numpy()
    assert got.dtype == np.uint8 and np.array_equal(got, c["ref"])       # incl. every integer-boundary colour triple
    with pytest.raises(v2v.V2VError):
        v2v.bgr_to_gray(torch.zeros((4, 4, 3), dtype=torch.uint8))


@pytest.mark.parametrize("noise", ["none", "philox"])
@pytest.mark.parametrize("shape", [(21, 64, 96, 4, True), (11, 30, 34, 2, False), (21, 480, 640, 1, True)])
def test_fused_pause_gather_and_degrade(cuda_device, noise, shape):
    """frames_to_voxel(raw stack, frame_index, value_map) == frames_to_voxel(gathered, degraded frames) == oracle."""
    import v2v_b200 as v2v
    n, h, w, B, first = shape
    kinds = ["hdr", "ldr", "hdr", "ldr"]
    raws, idxs, maps, prepared = [], [], [], []
    m_raw = 0
    np.random.seed(3)
    for b in range(B):
        idx, cnt = v2v.sample_pause_indices(n, 0.3, 0.6)
        raw = synth_video("walk", n, h, w, 40 + b)[:cnt]               # only `cnt` raw frames are decoded
        scale = np.random.uniform(1, 3) if kinds[b] == "hdr" else np.random.uniform(0.3, 1)
        vmap = v2v.degrade_value_map(kinds[b], scale)
        deg = [np.clip((im - 127.5) * scale + 127.5, 0, 255).astype(np.uint8) for im in raw]      # data/v2v_datasets.py:473-483
        prepared.append(np.stack([deg[i] for i in idx]))                                           # :311
        raws.append(raw), idxs.append(idx), maps.append(vmap)
        m_raw = max(m_raw, cnt)
    stack = np.zeros((B, m_raw, h, w), dtype=np.uint8)
    for b in range(B):
        stack[b, :raws[b].shape[0]] = raws[b]
    kw = dict(num_bins=5, frame_out="frames+first" if first else "frames", with_stats=True, noise=noise)
    if noise == "philox":
        kw.update(base_noise_std=0.05, hot_pixel_fraction=0.001, hot_pixel_std=2.0, seed=5)
    pos, neg = np.linspace(0.15, 0.4, B), np.linspace(0.3, 0.2, B)
    u0 = np.random.Generator(np.random.PCG64(2)).random((B, h, w))
    if noise == "none":
        kw.update(u0=u0)
    fused = v2v.frames_to_voxel(torch.from_numpy(stack).to(cuda_device), pos, neg, frame_index=np.stack(idxs),
                                value_map=np.stack(maps), **kw)
    plain = v2v.frames_to_voxel(torch.from_numpy(np.stack(prepared)).to(cuda_device), pos, neg, **kw)
    assert torch.equal(fused.voxel, plain.voxel) and torch.equal(fused.frames, plain.frames) and torch.equal(fused.stats, plain.stats)
    step = 5
    want_frames = np.stack(prepared)[:, (0 if first else step)::step].astype(np.float32) / np.float32(255)
    assert np.array_equal(fused.frames[:, :, 0].cpu().numpy(), want_frames)
    if noise == "none":
        for b in range(B):
            ref = orc.esim_video_to_voxel(prepared[b], float(pos[b]), float(neg[b]), 0.0, u0[b], np.zeros((h, w)),
                                          np.zeros((n - 1, h, w)), False)
            assert np.array_equal(fused.voxel[b].reshape(n - 1, h, w).cpu().numpy().astype(np.float64), ref)


def test_frame_index_is_clamped_and_validated(cuda_device):
    import v2v_b200 as v2v
    vid = synth_video("walk", 6, 16, 16, 1)
    fr = torch.from_numpy(vid).to(cuda_device)
    a = v2v.frames_to_voxel(fr, 0.2, 0.2, num_bins=5, frame_index=np.array([0, 1, 2, 3, 4, 99], dtype=np.int32))
    b = v2v.frames_to_voxel(fr, 0.2, 0.2, num_bins=5, frame_index=np.array([0, 1, 2, 3, 4, 5], dtype=np.int32))
    assert torch.equal(a.voxel, b.voxel)
    with pytest.raises(AssertionError):
        v2v.frames_to_voxel(fr, 0.2, 0.2, num_bins=5, frame_index=np.arange(5, dtype=np.int32))      # (N-1) % 5 != 0
    with pytest.raises(ValueError):
        v2v.frames_to_voxel(fr, 0.2, 0.2, num_bins=5, value_map=np.zeros(17, dtype=np.uint8))


@pytest.mark.parametrize("fast", ["1", "0"])
@pytest.mark.parametrize("noise", ["none", "philox"])
def test_v2e_fused_pause_gather_and_degrade(cuda_device, monkeypatch, fast, noise):
    """v2e kernels (throughput and generic, shot pre-pass, field dump): raw stack + frame_index + value_map == the same
    run on gathered, degraded frames, bit for bit."""
    import v2v_b200 as v2v
    from v2v_b200.v2e import frames_to_voxel_v2e
    if fast == "1":
        monkeypatch.setenv("V2V_V2E_FAST", "1")
    else:
        monkeypatch.setenv("V2V_V2E_GENERIC", "1")
    n, h, w, B = 13, 40, 64, 2
    raws, idxs, maps, prepared, m_raw = [], [], [], [], 0
    np.random.seed(11)
    for b in range(B):
        idx, cnt = v2v.sample_pause_indices(n, 0.3, 0.6)
        raw = synth_video("walk", n, h, w, 70 + b)[:cnt]
        scale = np.random.uniform(1, 3) if b == 0 else np.random.uniform(0.3, 1)
        vmap = v2v.degrade_value_map("hdr" if b == 0 else "ldr", scale)
        prepared.append(np.stack([vmap[raw[i]] for i in idx]))
        raws.append(raw), idxs.append(idx), maps.append(vmap)
        m_raw = max(m_raw, cnt)
    stack = np.zeros((B, m_raw, h, w), dtype=np.uint8)
    for b in range(B):
        stack[b, :raws[b].shape[0]] = raws[b]
    g = np.random.Generator(np.random.PCG64(6))
    pos = np.clip(g.normal(0.2, 0.05, (B, h, w)), 0.01, None)
    neg = np.clip(g.normal(0.2, 0.05, (B, h, w)), 0.01, None)
    nrate = np.exp(np.log(10) * 0.1 * g.standard_normal((B, h, w)).astype(np.float32)).astype(np.float32)
    kw = dict(fps=24, num_bins=3, frames_per_bin=2, cutoff_hz=30.0, leak_rate_hz=0.1, leak_jitter_fraction=0.1, noise_rate=nrate,
              with_stats=True, noise=noise)
    if noise == "philox":
        kw.update(shot_noise_rate_hz=20.0, seed=4, return_fields=True)
    fused = frames_to_voxel_v2e(torch.from_numpy(stack).to(cuda_device), pos, neg, frame_index=np.stack(idxs), value_map=np.stack(maps), **kw)
    plain = frames_to_voxel_v2e(torch.from_numpy(np.stack(prepared)).to(cuda_device), pos, neg, **kw)
    assert torch.equal(fused["voxel"], plain["voxel"]) and torch.equal(fused["stats"], plain["stats"])
    assert int(fused["stats"].sum()) > 0
    if noise == "philox":
        assert torch.equal(fused["shot_scales"], plain["shot_scales"])
        for k in ("leak_randn", "pos_shot", "neg_shot"):
            assert torch.equal(fused["fields"][k], plain["fields"][k])
