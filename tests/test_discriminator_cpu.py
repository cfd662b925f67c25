"""Discriminator (caller of conv2d_gradfix in the training step, BASELINE config 4) against the fixture written by the REAL
reference (oracle/make_golden_discriminator.py -> tests/golden/discriminator.npz): same state-dict names, same logits, and
the same parameter gradients of a D step with the R1 penalty (the double backward through every convolution)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from helpers import upfirdn2d_ref_on_cpu
from oracle import ref_generator
from oracle.make_golden_discriminator import CONFIGS, d_step

load_pkg()
disc = importlib.import_module('pgpp_b200.training.discriminator')
upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
NPZ = os.path.join(GOLDEN, 'discriminator.npz')


@pytest.mark.parametrize('name', list(CONFIGS))
def test_logits_and_r1_gradients_match_the_reference(name):
    g = np.load(NPZ)
    D = disc.Discriminator(**CONFIGS[name]).train().requires_grad_(True)
    assert sorted(f'{k}:{tuple(v.shape)}' for k, v in D.state_dict().items()) == [str(s) for s in g[f'{name}/state_names']]
    ref_generator.name_seeded_init(list(D.named_parameters()) + [(n, b) for n, b in D.named_buffers() if 'resample_filter' not in n])
    img, c = torch.from_numpy(g[f'{name}/img']), torch.from_numpy(g[f'{name}/c'])
    with upfirdn2d_ref_on_cpu(upf):
        logits, loss, grads = _step(D, img, c)
    assert torch.allclose(logits, torch.from_numpy(g[f'{name}/logits']), rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - float(g[f'{name}/loss'])) <= 1e-5 * max(1.0, abs(float(g[f'{name}/loss'])))
    want = {k[len(f'{name}/grad/'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f'{name}/grad/')}
    assert set(grads) == set(want)
    for k, v in grads.items():
        scale = max(want[k].abs().max().item(), 1e-6)
        assert (v - want[k]).abs().max().item() <= 2e-4 * scale, (k, (v - want[k]).abs().max().item(), scale)


def _step(D, img, c):
    """the fixture script's d_step on the package's discriminator through its explicit plain-PyTorch route (CPU tensors)"""
    class Wrapped(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.inner = D

        def forward(self, i, cc):
            return self.inner(i, cc, fused=False, impl='ref')

        def parameters(self, recurse=True):
            return self.inner.parameters(recurse)

        def named_parameters(self, *a, **k):
            return self.inner.named_parameters(*a, **k)
    return d_step(Wrapped(), img, c)


def test_freeze_d_and_minibatch_std():
    D = disc.Discriminator(c_dim=0, img_resolution=16, img_channels=3, channel_base=128, channel_max=16, block_kwargs=dict(freeze_layers=2))
    # frozen layers keep their tensors as buffers (networks.py:153-162): same state-dict names, not in parameters(), so the
    # training loop's `module.requires_grad_(True)` (training_loop_fullbody.py:613) cannot un-freeze them
    names = {n for n, _ in D.named_parameters()}
    frozen = [n for n, _ in D.named_buffers() if n.endswith('.weight') or n.endswith('.bias')]
    assert sorted(frozen) == ['b16.conv0.bias', 'b16.conv0.weight', 'b16.fromrgb.bias', 'b16.fromrgb.weight'] and not names & set(frozen)
    assert all(k in D.state_dict() for k in frozen)
    D.requires_grad_(True)
    assert not any(getattr(D.b16.conv0, k).requires_grad for k in ('weight', 'bias'))
    x = torch.randn(8, 4, 4, 4)
    y = disc.MinibatchStdLayer(group_size=4)(x)
    assert y.shape == (8, 5, 4, 4) and torch.equal(y[:, :4], x)
    grp = x.reshape(4, 2, 4, 4, 4)
    want = ((grp - grp.mean(0)).square().mean(0) + 1e-8).sqrt().mean(dim=[1, 2, 3])
    assert torch.allclose(y[:2, 4, 0, 0], want) and torch.allclose(y[2:4, 4, 0, 0], want)
