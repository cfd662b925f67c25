"""GPU parity of the discriminator / R1 training path (BASELINE configs[4]) on the sm_100a kernels: the fixture written by the REAL
reference (oracle/make_golden_discriminator.py -> tests/golden/discriminator.npz) replayed on cuda:0.  Logits, the loss and EVERY
parameter gradient of a D step with the R1 penalty -- forward, data gradient, weight gradient and the double backward of every
convolution run through pgpp_conv2d_igemm / pgpp_conv2d_wgrad, the FIR passes through pgpp_upfirdn2d, the activations through
pgpp_bias_act -- in both fp32-parity modes (`bf16x2` is what bench.py runs)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from oracle import ref_generator
from oracle.make_golden_discriminator import CONFIGS, d_step

pytestmark = pytest.mark.gpu
load_pkg()
disc = importlib.import_module('pgpp_b200.training.discriminator')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
NPZ = os.path.join(GOLDEN, 'discriminator.npz')
DEV = 'cuda:0'

# tolerance of a whole D step (about ten layers deep, R1 double backward on top): logits relative to max|logits|, every gradient
# relative to its own max-abs.  bf16x2 carries 16 significand bits per operand (per-layer rel-L2 <= 8e-5, tests/test_gpu_c_conv.py).
# measured on B200: logits <= 3e-5, worst gradient <= 1.6e-5 (bf16x2) / 6e-6 (bf16x3)
TOL = {'bf16x3': dict(logits=1e-4, grad=1e-4), 'bf16x2': dict(logits=1e-4, grad=1e-4)}


@pytest.fixture(autouse=True)
def _restore():
    old = cg.fp32_precision
    yield
    cg.fp32_precision = old


def _build(name):
    D = disc.Discriminator(**CONFIGS[name]).train().requires_grad_(True)
    ref_generator.name_seeded_init(list(D.named_parameters()) + [(n, b) for n, b in D.named_buffers() if 'resample_filter' not in n])
    return D.to(DEV)


@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3'])
@pytest.mark.parametrize('name', list(CONFIGS))
def test_d_step_with_r1_matches_the_reference_on_the_gpu_kernels(name, prec):
    g = np.load(NPZ)
    cg.fp32_precision = prec
    D = _build(name)
    img, c = torch.from_numpy(g[f'{name}/img']).to(DEV), torch.from_numpy(g[f'{name}/c']).to(DEV)
    before = custom_ops.launch_count()
    logits, loss, grads = d_step(D, img, c)
    assert custom_ops.launch_count() - before >= 50          # forward + R1 grad + double backward on the native kernels
    want_logits = torch.from_numpy(g[f'{name}/logits'])
    tol = TOL[prec]
    err = (logits.cpu() - want_logits).abs().max().item() / max(want_logits.abs().max().item(), 1e-6)
    assert err <= tol['logits'], ('logits', err)
    assert abs(float(loss) - float(g[f'{name}/loss'])) <= 10 * tol['logits'] * max(1.0, abs(float(g[f'{name}/loss'])))
    want = {k[len(f'{name}/grad/'):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(f'{name}/grad/')}
    assert set(grads) == set(want)
    worst = ('', 0.0)
    for k, v in grads.items():
        scale = max(want[k].abs().max().item(), 1e-6)
        e = (v.cpu() - want[k]).abs().max().item() / scale
        if e > worst[1]:
            worst = (k, e)
    print(f'{name} {prec}: logits err {err:.2e}, worst gradient {worst[0]} {worst[1]:.2e}')
    assert worst[1] <= tol['grad'], worst


@pytest.mark.parametrize('name', list(CONFIGS))
def test_inference_route_equals_training_route(name):
    """under no_grad the layers take the fused single-launch route; with gradients the conv2d_resample -> conv2d_gradfix route"""
    g = np.load(NPZ)
    D = _build(name)
    img, c = torch.from_numpy(g[f'{name}/img']).to(DEV), torch.from_numpy(g[f'{name}/c']).to(DEV)
    with torch.no_grad():
        a = D(img, c)
    b = D(img, c)
    assert b.requires_grad and not a.requires_grad
    want = torch.from_numpy(g[f'{name}/logits'])
    scale = want.abs().max().item()
    assert (a.cpu() - want).abs().max().item() <= 5e-4 * scale and (b.detach().cpu() - want).abs().max().item() <= 5e-4 * scale


def test_mixed_precision_blocks_stay_close_to_fp32():
    """num_fp16_res (networks.py:634,647): the fp16 blocks keep logits within fp16 accuracy of the fp32 network"""
    cfg = dict(CONFIGS['resnet_cond'])
    g = np.load(NPZ)
    img, c = torch.from_numpy(g['resnet_cond/img']).to(DEV), torch.from_numpy(g['resnet_cond/c']).to(DEV)
    outs = []
    for nfp16 in (0, 2):
        D = disc.Discriminator(**cfg, num_fp16_res=nfp16).train().requires_grad_(True)
        ref_generator.name_seeded_init(list(D.named_parameters()) + [(n, b) for n, b in D.named_buffers() if 'resample_filter' not in n])
        D = D.to(DEV)
        logits, loss, grads = d_step(D, img, c)
        outs.append((logits, grads))
    scale = outs[0][0].abs().max().item()
    assert (outs[0][0] - outs[1][0]).abs().max().item() <= 2e-2 * scale
    for k in outs[0][1]:
        s = max(outs[0][1][k].abs().max().item(), 1e-6)
        assert (outs[0][1][k] - outs[1][1][k].float()).abs().max().item() <= 0.1 * s, k
