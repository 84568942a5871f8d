"""Property tests of the oracle (SURVEY §8 oracle plan, item 5): invariants of the reference's algorithm that hold for any
input, checked on random small cases with hypothesis.  They guard the checker itself; the GPU twins are in
test_properties_gpu.py."""
import numpy as np
from hypothesis import given, settings, strategies as st

import v2v_oracle as orc

SET = dict(max_examples=40, deadline=None, derandomize=True)


def _video(rs, n, h, w):
    base = rs.randint(0, 256, (h, w)).astype(np.int64)
    steps = rs.randint(-40, 41, (n, h, w))
    steps[0] = 0
    return np.clip(base[None] + np.cumsum(steps, 0), 0, 255).astype(np.uint8)


@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(2, 9), h=st.integers(1, 7), w=st.integers(1, 9),
       pos=st.floats(0.05, 1.5), gap=st.floats(1.0, 1.5), std=st.floats(0.0, 0.3))
@settings(**SET)
def test_esim_integrator_conserves_log_intensity(seed, n, h, w, pos, gap, std):
    """sum_i (pe_i*pos - ne_i*neg) + pot_end == pot_0 + L_end - L_0 + sum noise (data/v2v_core_esim.py:42-58), the residual
    potential stays inside (-neg, pos), and a pixel-interval never holds events of both signs."""
    rs = np.random.RandomState(seed)
    neg = pos * gap
    video = _video(rs, n, h, w)
    u0, hot, g = orc.esim_draw_randomness(n, h, w, 0.2, 0.5, rs)
    out, pot = orc.esim_video_to_voxel(video, pos, neg, std, u0, hot, g, False, return_state=True)
    lut = orc.esim_log_lut()
    L = lut[video]
    pot0 = u0 * (pos + neg) - neg
    injected = (L[-1] - L[0]) + std * g.sum(0) + (n - 1) * hot
    removed = np.maximum(out, 0).sum(0) * pos - np.maximum(-out, 0).sum(0) * neg
    assert np.allclose(pot0 + injected - removed, pot, rtol=0, atol=1e-9 * (n + np.abs(out).sum(0).max()))
    assert np.all(pot < pos) and np.all(pot > -neg)
    assert np.all(out == np.rint(out))                        # integer counts (noise internal)


@given(seed=st.integers(0, 2 ** 31 - 1), n=st.integers(3, 13), bins=st.sampled_from([1, 2, 3, 4, 6]), fpb=st.sampled_from([1, 2, 3]))
@settings(**SET)
def test_bin_accumulate_preserves_totals(seed, n, bins, fpb):
    rs = np.random.RandomState(seed)
    t = max(1, n // 2)
    x = rs.randint(-5, 6, (t * bins * fpb, 3, 4)).astype(np.float64)
    v = orc.bin_accumulate(x, bins, fpb)
    assert v.shape == (t, bins, 3, 4)
    assert np.array_equal(v.sum((0, 1)), x.sum(0))
    assert np.array_equal(v[0, 0], x[:fpb].sum(0))


@given(seed=st.integers(0, 2 ** 31 - 1), ne=st.integers(0, 400), h=st.integers(1, 12), w=st.integers(1, 14),
       bins=st.sampled_from([1, 2, 5, 15]), f32=st.booleans())
@settings(**SET)
def test_make_voxel_conserves_signed_counts(seed, ne, h, w, bins, f32):
    """Discrete mode: every event lands in exactly one cell (sum voxel == sum (2p-1)), per pixel too.  Interpolated mode:
    the two temporal taps of an event sum to one as long as t_norm <= bins-1, which (:76-77) always holds
    (data/testh5.py:60-90)."""
    rs = np.random.RandomState(seed)
    ts = np.sort(rs.rand(ne) * 0.05 + 3.0).astype(np.float32 if f32 else np.float64)
    xs = rs.randint(0, w, ne).astype(np.uint16)
    ys = rs.randint(0, h, ne).astype(np.uint16)
    ps = rs.randint(0, 2, ne).astype(np.uint8)
    pol = 2.0 * ps - 1.0
    per_pixel = np.zeros((h, w))
    np.add.at(per_pixel, (ys, xs), pol)
    d = orc.make_voxel(ts, xs, ys, ps, bins, h, w, False)
    assert d.shape == (bins, h, w)
    assert np.array_equal(d.sum(0), per_pixel)
    if bins > 1:
        i = orc.make_voxel(ts, xs, ys, ps, bins, h, w, True)
        assert np.allclose(i.sum(0), per_pixel, rtol=0, atol=1e-9 * max(1, ne))


@given(seed=st.integers(0, 2 ** 31 - 1), ne=st.integers(1, 300), bins=st.sampled_from([1, 3, 5]))
@settings(**SET)
def test_torch_voxel_discrete_conserves_weights(seed, ne, bins):
    rs = np.random.RandomState(seed)
    h, w = 9, 11
    ts = np.sort(rs.rand(ne)).astype(np.float32)
    ts -= ts[0]
    xs = rs.randint(0, w, ne).astype(np.float32)
    ys = rs.randint(0, h, ne).astype(np.float32)
    ps = (2.0 * rs.randint(0, 2, ne) - 1.0).astype(np.float32)
    v = orc.events_to_voxel_f32(xs, ys, ts, ps, bins, (h, w), False)
    assert v.shape == (bins, h, w)
    assert float(v.sum()) == float(ps.sum())                  # small integers: float32 sums are exact
