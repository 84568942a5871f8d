"""The mixins composed with the reference's REAL dataset classes (the drop-in seam is the yaml
``class_name`` resolved to ``Dataset(data_path, configs)``, data/data_interface.py:6-20).

Needs the reference checkout (``/root/reference`` in the build container; skipped where it is absent — the GPU box gets
only this repo).  ``class X(ImgsToVoxelsMixin, WebvidDatasetV2)`` / ``class Y(MakeVoxelMixin, TestH5Dataset)`` are built
exactly as a maintainer would; the test checks that the override is the method the reference's own ``__getitem__`` /
caller reaches, with the reference's arguments, and that the item equals the unmodified class's item under the same
``np.random`` seed.  With a CUDA device the override runs the kernels; without one (the build container) the single GPU
call underneath is replaced by the CPU oracle, so what is checked here is the composition and the host logic around it
(parameter draw order, reshape / bin sum, packing), while tests/test_esim_gpu.py / test_scatter_gpu.py check the kernels
against goldens recorded from these same reference classes.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

import v2v_oracle as orc
from conftest import synth_video

REF = os.environ.get("V2V_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "data")), reason="reference checkout not present")


class _Node:
    """An h5 dataset stand-in: array access, ``[()]`` and ``.attrs`` (the reference only reads)."""

    def __init__(self, arr, **attrs):
        self.arr, self.attrs, self.shape = arr, attrs, arr.shape

    def __getitem__(self, k):
        return self.arr if isinstance(k, tuple) and k == () else self.arr[k]


class _FakeH5:
    store = {}

    def __init__(self, path, mode="r"):
        self.tree = _FakeH5.store[path]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, key):
        return self.tree[key]


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    for m in ("h5py", "matplotlib", "matplotlib.pyplot", "ffmpeg"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    evb = types.ModuleType("event_voxel_builder")
    evb.EventVoxelBuilder = object
    sys.modules.setdefault("event_voxel_builder", evb)
    import data.testh5 as testh5
    import data.v2v_datasets as dsets
    yield types.SimpleNamespace(testh5=testh5, dsets=dsets)
    sys.path.remove(REF)


class _OracleEmulator:
    """CPU stand-in for v2v_b200.EventEmulator(rng="numpy") when there is no GPU: same constructor, same draws."""

    def __init__(self, pos_thres, neg_thres, base_noise_std, hot_pixel_fraction, hot_pixel_std, put_noise_external, seed=None,
                 *, rng=None, device=None):
        self.a = (pos_thres, neg_thres, base_noise_std, hot_pixel_fraction, hot_pixel_std, put_noise_external)
        assert rng == "numpy"

    def video_to_voxel(self, video):
        pos, neg, std, hf, hs, ext = self.a
        n, h, w = video.shape
        u0, hot, g = orc.esim_draw_randomness(n, h, w, hf, hs)
        return orc.esim_video_to_voxel(video, pos, neg, std, u0, hot, g, ext)


@pytest.mark.parametrize("cfg", [dict(), dict(frames_per_bin=2, scale_noise_strength=True), dict(put_noise_external=True)])
def test_imgs_to_voxels_mixin_on_webvid_dataset(ref, monkeypatch, cfg):
    import v2v_b200.datasets as mine
    if not torch.cuda.is_available():
        monkeypatch.setattr(mine, "EventEmulator", _OracleEmulator)

    class WebvidDatasetV2B200(mine.ImgsToVoxelsMixin, ref.dsets.WebvidDatasetV2):
        v2v_rng = "numpy"

    assert WebvidDatasetV2B200.imgs_to_voxels is mine.ImgsToVoxelsMixin.imgs_to_voxels
    assert WebvidDatasetV2B200.__getitem__ is ref.dsets.WebvidDatasetV2.__getitem__       # everything else is inherited
    configs = dict(sequence_length=2, num_bins=5, base_noise_std_range=[0, 0.1], hot_pixel_std_range=[0, 10], **cfg)
    new = WebvidDatasetV2B200.__new__(WebvidDatasetV2B200)
    old = ref.dsets.WebvidDatasetV2.__new__(ref.dsets.WebvidDatasetV2)
    new.load_configs(configs), old.load_configs(configs)
    fpb = configs.get("frames_per_bin", 1)
    imgs = synth_video("walk", 2 * 5 * fpb + 1, 24, 32, 5)
    for fixed in ((None, None), (0.31, 0.22)):
        new.use_fixed_thresholds = old.use_fixed_thresholds = fixed[0] is not None
        np.random.seed(17)
        p_ref, v_ref = old.imgs_to_voxels(imgs, 5, fpb, 24, *fixed)
        np.random.seed(17)
        p_new, v_new = new.imgs_to_voxels(imgs, 5, fpb, 24, *fixed)
        assert p_new == p_ref
        assert v_new.shape == v_ref.shape and v_new.dtype == v_ref.dtype and np.array_equal(v_new, v_ref)


@pytest.mark.parametrize("interp", [False, True])
def test_make_voxel_mixin_on_testh5_dataset(ref, monkeypatch, interp):
    import v2v_b200.events as mine
    if not torch.cuda.is_available():
        monkeypatch.setattr(mine, "make_voxel", lambda evs, bins, H, W, ib, device=None: orc.make_voxel(*evs, bins, H, W, ib))
    g = np.random.Generator(np.random.PCG64(8))
    H, W, n_img, ne = 20, 28, 5, 4000
    ts = np.sort(g.random(ne)) * 0.2 + 3.0
    idx = np.linspace(0, ne, n_img).astype(np.int64)
    idx[2] = idx[1]                                        # an empty window
    images = {f"image{k:09d}": _Node(g.integers(0, 256, (H, W)).astype(np.uint8), event_idx=int(idx[k])) for k in range(n_img)}
    _FakeH5.store["/fake/hqf_x.h5"] = {"images": images, "events/ts": ts, "events/xs": g.integers(0, W, ne).astype(np.uint16),
                                       "events/ys": g.integers(0, H, ne).astype(np.uint16),
                                       "events/ps": (g.random(ne) < 0.5).astype(np.uint8)}
    monkeypatch.setattr(ref.testh5.h5py, "File", _FakeH5, raising=False)

    class TestH5DatasetB200(mine.MakeVoxelMixin, ref.testh5.TestH5Dataset):
        pass

    TestH5DatasetB200.__test__ = False
    assert TestH5DatasetB200.make_voxel is mine.MakeVoxelMixin.make_voxel
    configs = dict(sequence_length=4, num_bins=5, interpolate_bins=interp, output_additional_evs=True)
    new = TestH5DatasetB200("/fake/hqf_x.h5", configs)     # the reference's own constructor and __getitem__
    old = ref.testh5.TestH5Dataset("/fake/hqf_x.h5", configs)
    a, b = new[0], old[0]
    assert a["events"].shape == b["events"].shape == (5, 5, H, W)
    assert torch.equal(a["frame"], b["frame"]) and a["sequence_name"] == b["sequence_name"]
    assert torch.allclose(a["events"], b["events"], rtol=1e-5, atol=1e-6)
    assert float(a["events"][2].abs().sum()) == 0.0       # (first slot is the additional window; window 1 -> 2 is empty)
