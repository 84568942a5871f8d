"""GPU parity of the consumer-side helpers: normalize_batch_voxel against the reference's own expression
(model/train_utils.py:147-166, torch ops) and the per-bin sums / value histogram against NumPy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_normalize_batch_voxel(voxel):
    """model/train_utils.py:147-166, statement for statement (torch; runs on the tensor's device)."""
    assert len(voxel.shape) == 5
    B, T, C, H, W = voxel.shape
    voxel_flat = voxel.reshape((B, -1))
    max_k = int(0.99 * voxel_flat.shape[1])
    min_k = int(0.01 * voxel_flat.shape[1])
    pos_max = torch.kthvalue(voxel_flat, max_k, dim=1).values
    neg_max = -torch.kthvalue(voxel_flat, min_k, dim=1).values
    pos_max = torch.clamp(pos_max.reshape((B, 1, 1, 1, 1)), min=1)
    neg_max = torch.clamp(neg_max.reshape((B, 1, 1, 1, 1)), min=1)
    return torch.where(voxel > 0, voxel / pos_max, voxel / neg_max)


@pytest.mark.parametrize("kind", ["sparse", "dense", "heavy", "noisy"])
def test_normalize_batch_voxel_matches_reference(cuda_device, kind):
    import v2v_b200 as v2v
    g = np.random.Generator(np.random.PCG64(3))
    shape = (3, 4, 5, 36, 44)
    if kind == "sparse":          # < 1 % events per polarity: both factors clamp to 1 -> identity, no second pass
        v = (g.random(shape) < 0.004).astype(np.float32) - (g.random(shape) < 0.004).astype(np.float32)
    elif kind == "dense":         # 99th percentile is a single event
        v = g.integers(-1, 2, shape).astype(np.float32)
    elif kind == "heavy":         # clips with different, larger order statistics
        v = np.stack([g.integers(-s, s + 1, shape[1:]) for s in (2, 9, 40)]).astype(np.float32)
    else:                         # non-integer voxels (external noise): the kthvalue fallback
        v = (g.integers(-3, 4, shape) + 0.3 * g.standard_normal(shape)).astype(np.float32)
    t = torch.from_numpy(v).to(cuda_device)
    ref = ref_normalize_batch_voxel(t)
    got = v2v.normalize_batch_voxel(t)
    assert torch.equal(got, ref)
    assert torch.equal(t, torch.from_numpy(v).to(cuda_device))            # input untouched
    v2v.normalize_batch_voxel(t, inplace=True)
    assert torch.equal(t, ref)


def test_value_hist_and_bin_sums(cuda_device):
    import v2v_b200 as v2v
    g = np.random.Generator(np.random.PCG64(5))
    v = g.integers(-300, 301, (2, 3, 5, 20, 28)).astype(np.float32)
    v[0, 0, 0, 0, :5] = [0.5, -0.25, 1e9, -1e9, 2.5]
    t = torch.from_numpy(v).to(cuda_device)
    h = v2v.voxel_value_hist(t).cpu().numpy()
    for b in range(2):
        x = v[b].ravel()
        ok = (x == np.rint(x)) & (np.abs(x) <= 255)
        assert h[b, 511] == (~ok).sum()
        assert np.array_equal(h[b, :511], np.bincount(x[ok].astype(np.int64) + 255, minlength=511))
    s = v2v.bin_abs_sums(t).cpu().numpy()
    assert np.array_equal(s, np.abs(v.astype(np.float64)).sum(axis=(0, 1, 3, 4)))
