"""NER-Net's quantization-layer scatter (model/nernet/representation_modules.py:143-168, 228-248) as a client of the library:
all bins in one launch, differentiable; checked against the reference's own statement sequence run with torch on the GPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _events(n, W, H, B, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, W, (n,), generator=g)
    y = torch.randint(0, H, (n,), generator=g)
    p = torch.randint(0, 2, (n,), generator=g)
    b = torch.sort(torch.randint(0, B, (n,), generator=g)).values
    t = torch.rand(n, generator=g)
    return x, y, t, p, b


class _ValueLayer(torch.nn.Module):            # the shape of the reference's ValueLayer MLP: 1 -> 30 -> 30 -> 1
    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Linear(1, 30), torch.nn.LeakyReLU(0.1), torch.nn.Linear(30, 30), torch.nn.LeakyReLU(0.1),
                                       torch.nn.Linear(30, 1))

    def forward(self, x):
        return self.net(x[..., None]).squeeze(-1)


@pytest.mark.parametrize("two_pol", [True, False])
def test_quantization_layer_scatter_matches_reference_statements(cuda_device, two_pol):
    import v2v_b200 as v2v
    C, H, W, B, n = 5, 24, 32, 3, 20000
    x, y, t, p, b = (a.to(cuda_device) for a in _events(n, W, H, B, 3))
    torch.manual_seed(0)
    layer = _ValueLayer().to(cuda_device)
    P = 2 if two_pol else 1
    num_voxels = int(P * C * H * W * B)
    # :137-140 (one polarity, signed t) / the two-polarity layer further down: index of every event without the bin term
    idx_before_bins = (x + W * y + 0 + W * H * C * p * (1 if two_pol else 0) + W * H * C * P * b).long()

    def reference():
        vox = torch.zeros(num_voxels, device=cuda_device)
        for i_bin in range(C):                                                  # :143-168
            values = t * layer(t - i_bin / (C - 1))
            idx = torch.clamp((idx_before_bins + W * H * i_bin).long(), max=vox.shape[0] - 1)
            vox.put_(idx, values, accumulate=True)
        return vox

    def ours():
        vox = torch.zeros(num_voxels, device=cuda_device)
        values = torch.stack([t * layer(t - i_bin / (C - 1)) for i_bin in range(C)])
        return v2v.put_accumulate_bins(vox, idx_before_bins, values, W * H)

    ref = reference()
    got = ours()
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
    # gradients of the MLP through the scatter: d/dtheta of sum(vox * weights)
    wts = torch.randn(num_voxels, device=cuda_device)
    layer.zero_grad()
    (reference() * wts).sum().backward()
    gref = [q.grad.clone() for q in layer.parameters()]
    layer.zero_grad()
    (ours() * wts).sum().backward()
    for a, q in zip(gref, layer.parameters()):
        assert torch.allclose(q.grad, a, rtol=1e-4, atol=1e-5)


def test_put_accumulate_single_bin_and_clamp(cuda_device):
    import v2v_b200 as v2v
    g = torch.Generator().manual_seed(1)
    idx = torch.randint(0, 1200, (5000,), generator=g).to(cuda_device)          # some indices beyond the end: clamped like :166
    val = torch.randn(5000, generator=g).to(cuda_device)
    ref = torch.zeros(1000, device=cuda_device)
    ref.put_(torch.clamp(idx, max=999), val, accumulate=True)
    got = v2v.put_accumulate(torch.zeros(1000, device=cuda_device), idx, val)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
