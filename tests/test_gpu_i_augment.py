"""AugmentPipe on the GPU: the deterministic (`debug_percentile`) runs of the reference fixture through this package's CUDA kernels
(upfirdn2d up/down-sampling with the 12-tap sym6 filter, grid_sample), and the R1 pattern through the pipeline."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from oracle.make_golden_augment import CONFIGS, RUNS

pytestmark = pytest.mark.gpu
load_pkg()
augment = importlib.import_module('pgpp_b200.training.augment')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'


@pytest.mark.parametrize('run', [r for r in RUNS if r[4] is not None], ids=[r[0] for r in RUNS if r[4] is not None])
def test_deterministic_runs_match_the_reference_fixture(run):
    name, cfg, p, seed, dbg, shape = run
    g = np.load(os.path.join(GOLDEN, 'augment.npz'))
    pipe = augment.AugmentPipe(**CONFIGS[cfg]).to(DEV).eval().requires_grad_(False)
    pipe.p.copy_(torch.as_tensor(p))
    before = custom_ops.launch_count()
    if 'noise' in CONFIGS[cfg]:
        pytest.skip('additive noise draws from the device RNG: not comparable across devices')
    y = pipe(torch.from_numpy(g[f'{name}/x']).to(DEV), debug_percentile=dbg)
    want = torch.from_numpy(g[f'{name}/y'])
    assert custom_ops.launch_count() - before >= 5              # 2 + 2 separable FIR passes and the grid sample ran on this package's kernels
    err = (y.cpu() - want).abs().max().item()
    assert err <= 2e-4 * max(1.0, want.abs().max().item()), err


def test_random_mode_and_r1_through_the_pipeline():
    torch.manual_seed(3)
    pipe = augment.AugmentPipe(**CONFIGS['bgc']).to(DEV)
    x = torch.randn(8, 3, 64, 64, device=DEV).clamp(-1, 1)
    y = pipe(x)
    assert y.shape == x.shape and torch.isfinite(y).all()
    s = torch.tensor(1.0, device=DEV, requires_grad=True)
    xi = (x * s).requires_grad_(True)
    torch.manual_seed(4)
    out = pipe(xi)
    gx, = torch.autograd.grad(out.square().sum(), [xi], create_graph=True)
    gs, = torch.autograd.grad(gx.square().sum(), [s])
    assert torch.isfinite(gs) and gs.item() > 0
