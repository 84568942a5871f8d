"""GPU parity of the voxel-space noise augmentation and of the cached-voxel dataset item against
outputs of the reference's own code (data/esim_dataset.py:7-46,84-153), recorded in tests/golden/augment.npz by
tests/golden/make_golden.py (which also asserts the oracle's restatement against the reference)."""
import random

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden("augment").names("noise_"))
def test_add_noise_numpy_rng_reproduces_reference(cuda_device, name):
    from v2v_b200 import augment
    c = golden("augment").case(name)
    np.random.seed(int(c["seed"]))
    got = augment.add_noise_to_voxel(torch.from_numpy(c["voxel"].copy()).to(cuda_device), float(c["noise_std"]),
                                     float(c["noise_fraction"]), bool(c["integer_noise"]), rng="numpy")
    assert np.array_equal(got.cpu().numpy(), c["ref"].astype(np.float32))


@pytest.mark.parametrize("name", golden("augment").names("hot_"))
def test_add_hot_pixels_reproduces_reference(cuda_device, name):
    from v2v_b200 import augment
    c = golden("augment").case(name)
    np.random.seed(int(c["np_seed"])), random.seed(int(c["py_seed"]))
    got = augment.add_hot_pixels_to_voxels(torch.from_numpy(c["voxels"].copy()).to(cuda_device), float(c["hot_pixel_std"]),
                                           float(c["max_hot_pixel_fraction"]), bool(c["integer_noise"]))
    # float64 atomics build the [H,W] map: pixels hit twice may differ in the last float64 bit before the float32 store
    assert np.allclose(got.cpu().numpy(), c["ref"], rtol=0, atol=1e-6)
    assert (got.cpu().numpy() != c["voxels"]).any()


@pytest.mark.parametrize("name", golden("augment").names("item_"))
def test_cached_sequence_item_reproduces_reference(cuda_device, name):
    """ESIMH5Dataset.__getitem__ (pause sequence + per-step noise + hot pixels) on the GPU == the reference's item."""
    from v2v_b200 import augment
    c = golden("augment").case(name)
    dev = cuda_device
    np.random.seed(int(c["np_seed"])), random.seed(int(c["py_seed"]))
    item = augment.cached_sequence_item(torch.from_numpy(c["frames"]).to(dev), torch.from_numpy(c["flow"]).to(dev),
                                        torch.from_numpy(c["events"]).to(dev), int(c["sequence_length"]),
                                        proba_pause_when_running=float(c["proba_pause_when_running"]),
                                        proba_pause_when_paused=float(c["proba_pause_when_paused"]),
                                        noise_std=float(c["noise_std"]), noise_fraction=float(c["noise_fraction"]),
                                        hot_pixel_std=float(c["hot_pixel_std"]),
                                        max_hot_pixel_fraction=float(c["max_hot_pixel_fraction"]),
                                        integer_noise=bool(c["integer_noise"]), rng="numpy")
    assert np.array_equal(item["frame"].cpu().numpy(), c["ref_frame"])
    assert np.array_equal(item["flow"].cpu().numpy(), c["ref_flow"])
    assert np.allclose(item["events"].cpu().numpy(), c["ref_events"], rtol=0, atol=1e-6)
    paused = c["src"] < 0
    assert float(item["flow"][torch.from_numpy(paused).to(dev)].abs().sum()) == 0.0


def test_philox_mode_statistics_and_fresh_default_streams(cuda_device):
    """Throughput mode: in-kernel noise with the requested std / fraction; two calls with the reference's signature (no
    seed) must not repeat the noise field (the reference draws fresh np.random noise on every call)."""
    from v2v_b200 import augment
    z = torch.zeros((4, 5, 128, 160), dtype=torch.float32, device=cuda_device)
    a = augment.add_noise_to_voxel(z.clone(), 0.5, 0.25)
    b = augment.add_noise_to_voxel(z.clone(), 0.5, 0.25)
    assert not torch.equal(a, b)
    nz = (a != 0).float().mean().item()
    assert abs(nz - 0.25) < 0.01
    assert abs(a[a != 0].std().item() - 0.5) < 0.01
    assert abs(torch.corrcoef(torch.stack([a.flatten(), b.flatten()]))[0, 1].item()) < 0.01
    c = augment.add_noise_to_voxel(z.clone(), 0.5, 0.25, seed=11, stream_id=3)
    d = augment.add_noise_to_voxel(z.clone(), 0.5, 0.25, seed=11, stream_id=3)
    assert torch.equal(c, d)                       # explicit seed and stream: reproducible
    i = augment.add_noise_to_voxel(z.clone(), 1.5, 1.0, integer_noise=True)
    assert torch.equal(i, i.round()) and abs(i.std().item() - 1.5) < 0.05
