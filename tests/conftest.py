import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))       # the oracle is test infrastructure only
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


class Golden:
    """Access to tests/golden/<group>.npz written by tests/golden/make_golden.py."""

    def __init__(self, group):
        self.z = np.load(os.path.join(GOLDEN, f"{group}.npz"))
        self.cases = [str(c) for c in self.z["__cases__"]]

    def case(self, name):
        pre = name + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def names(self, prefix=""):
        return [c for c in self.cases if c.startswith(prefix)]


_cache = {}


def golden(group):
    if group not in _cache:
        _cache[group] = Golden(group)
    return _cache[group]


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def synth_video(kind, n, h, w, seed):
    g = np.random.Generator(np.random.PCG64(seed))
    if kind == "walk":
        base = g.integers(0, 256, size=(h, w)).astype(np.int64)
        steps = g.integers(-6, 7, size=(n, h, w))
        steps[0] = 0
        return np.clip(base[None] + np.cumsum(steps, axis=0), 0, 255).astype(np.uint8)
    if kind == "iid":
        return g.integers(0, 256, size=(n, h, w), dtype=np.uint8)
    raise ValueError(kind)
