"""GPU parity of the voxel-space noise augmentation against the reference's own code, restated
inline from data/esim_dataset.py:7-46 (NumPy, same seeds)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref_add_noise(voxel, noise_std, noise_fraction, integer_noise):          # data/esim_dataset.py:33-46
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * noise_std ** 2)) / 2
        y = np.random.poisson(lam=lmb, size=voxel.shape)
        sign = 2 * np.random.randint(0, 2, size=voxel.shape) - 1
        noise = y * sign
    else:
        noise = noise_std * np.random.randn(*voxel.shape)
    if noise_fraction < 1.0:
        mask = np.random.rand(*voxel.shape) >= noise_fraction
        noise = np.where(mask, 0, noise)
    return voxel + noise


def ref_hot(voxels, hot_pixel_std, max_frac, integer_noise):                # data/esim_dataset.py:7-30
    T, C, H, W = voxels.shape
    frac = random.uniform(0, max_frac)
    num = int(frac * H * W)
    x = np.random.randint(0, W, num)
    y = np.random.randint(0, H, num)
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * hot_pixel_std ** 2)) / 2
        yy = np.random.poisson(lam=lmb, size=num)
        sign = 2 * np.random.randint(0, 2, size=num) - 1
        val = yy * sign
    else:
        val = np.random.randn(num)
        val *= hot_pixel_std
    noise = np.zeros((H, W))
    np.add.at(noise, (y, x), val)
    return voxels + noise[np.newaxis, np.newaxis, ...]


@pytest.mark.parametrize("integer_noise", [False, True])
def test_numpy_rng_reproduces_reference(cuda_device, integer_noise):
    from v2v_b200 import augment
    g = np.random.Generator(np.random.PCG64(4))
    vox = g.integers(-3, 4, size=(6, 5, 40, 56)).astype(np.float32)
    np.random.seed(21)
    ref = ref_add_noise(vox, 0.7, 0.1, integer_noise)
    np.random.seed(21)
    got = augment.add_noise_to_voxel(torch.from_numpy(vox.copy()).to(cuda_device), 0.7, 0.1, integer_noise, rng="numpy")
    assert np.array_equal(got.cpu().numpy(), ref.astype(np.float32))
    np.random.seed(22); random.seed(5)
    refh = ref_hot(vox.astype(np.float64), 2.0, 0.05, integer_noise)
    np.random.seed(22); random.seed(5)
    goth = augment.add_hot_pixels_to_voxels(torch.from_numpy(vox.copy()).to(cuda_device), 2.0, 0.05, integer_noise)
    assert np.allclose(goth.cpu().numpy(), refh.astype(np.float32), rtol=0, atol=1e-6)
    assert (goth.cpu().numpy() != vox).any()


@pytest.mark.parametrize("integer_noise", [False, True])
def test_philox_noise_statistics(cuda_device, integer_noise):
    from v2v_b200 import augment
    vox = torch.zeros((8, 5, 128, 128), dtype=torch.float32, device=cuda_device)
    std, frac = 1.3, 0.25
    a = augment.add_noise_to_voxel(vox.clone(), std, frac, integer_noise, rng="philox", seed=3)
    b = augment.add_noise_to_voxel(vox.clone(), std, frac, integer_noise, rng="philox", seed=3)
    c = augment.add_noise_to_voxel(vox.clone(), std, frac, integer_noise, rng="philox", seed=4)
    assert torch.equal(a, b) and not torch.equal(a, c)
    x = a.cpu().numpy().ravel()
    nz = x != 0
    if integer_noise:
        lmb = (-1 + np.sqrt(1 + 4 * std ** 2)) / 2
        p_nz = frac * (1 - np.exp(-lmb))
        assert abs(nz.mean() - p_nz) / p_nz < 0.03
        assert abs(np.abs(x[nz]).mean() - lmb / (1 - np.exp(-lmb))) < 0.03 and abs(x[nz].mean()) < 0.03
        assert np.all(x == np.round(x))
    else:
        assert abs(nz.mean() - frac) < 0.005
        assert abs(x[nz].std() - std) / std < 0.02 and abs(x[nz].mean()) < 0.02
