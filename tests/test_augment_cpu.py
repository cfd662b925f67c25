"""AugmentPipe against the fixture the REAL reference wrote (oracle/make_golden_augment.py ->
tests/golden/augment.npz): with the same seed on the same device the package's pipeline draws the same random tensors in the
same order and must reproduce the reference's images; `debug_percentile` runs are deterministic and device independent."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from helpers import upfirdn2d_ref_on_cpu
from oracle.make_golden_augment import CONFIGS, RUNS

load_pkg()
augment = importlib.import_module('pgpp_b200.training.augment')
upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
NPZ = os.path.join(GOLDEN, 'augment.npz')


@pytest.fixture(scope='module')
def golden():
    return dict(np.load(NPZ))


def test_buffers_match_the_reference(golden):
    pipe = augment.AugmentPipe()
    assert sorted(pipe.state_dict()) == ['Hz_fbank', 'Hz_geom', 'p']
    assert np.allclose(pipe.Hz_geom.numpy(), golden['Hz_geom'], rtol=0, atol=1e-7)
    assert np.allclose(pipe.Hz_fbank.numpy(), golden['Hz_fbank'], rtol=0, atol=1e-7)
    assert pipe.xint_max == 0.125 and pipe.imgfilter_bands == [1, 1, 1, 1] and float(pipe.p) == 1.0


@pytest.mark.parametrize('run', RUNS, ids=[r[0] for r in RUNS])
def test_reproduces_the_reference_images(golden, run):
    name, cfg, p, seed, dbg, shape = run
    pipe = augment.AugmentPipe(**CONFIGS[cfg]).eval().requires_grad_(False)
    pipe.p.copy_(torch.as_tensor(p))
    x = torch.from_numpy(golden[f'{name}/x'])
    assert tuple(x.shape) == shape
    torch.manual_seed(seed + 100)
    with upfirdn2d_ref_on_cpu(upf):
        y = pipe(x, debug_percentile=dbg)
    want = torch.from_numpy(golden[f'{name}/y'])
    assert y.shape == want.shape and y.dtype == torch.float32
    err = (y - want).abs().max().item()
    assert err <= 2e-5 * max(1.0, want.abs().max().item()), err


def test_identity_when_everything_is_off_and_gradients_flow():
    x = torch.randn(2, 3, 16, 16)
    assert augment.AugmentPipe()(x) is x
    pipe = augment.AugmentPipe(scale=1, rotate=1, brightness=1, saturation=1)
    xin = x.clone().requires_grad_(True)
    with upfirdn2d_ref_on_cpu(upf):
        y = pipe(xin)
    g, = torch.autograd.grad(y.square().sum(), [xin])
    assert g.shape == x.shape and torch.isfinite(g).all() and g.abs().sum() > 0
    with pytest.raises(ValueError):
        augment.AugmentPipe(brightness=1)(torch.randn(1, 2, 8, 8))
