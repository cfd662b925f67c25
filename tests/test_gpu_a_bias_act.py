"""GPU parity tests for bias_act (through the C ABI: pgpp_bias_act)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from helpers import max_abs, rel_l2, t
from oracle import ref_ops

pytestmark = pytest.mark.gpu
load_pkg()
bias_act = importlib.import_module('pgpp_b200.torch_utils.ops.bias_act')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'


def _opt(v):
    return None if v == 'None' else float(v)


def test_golden_forward_backward_double_backward():
    g = np.load(os.path.join(GOLDEN, 'bias_act.npz'))
    n = len([k for k in g.files if k.endswith('_meta')])
    before = custom_ops.launch_count()
    for i in range(n):
        act, alpha, gain, clamp = g[f'case{i}_meta']
        act = str(act); alpha, gain, clamp = _opt(alpha), _opt(gain), _opt(clamp)
        x = t(g['x']).to(DEV).requires_grad_(True)
        b = t(g['b']).to(DEV).requires_grad_(True)
        y = bias_act.bias_act(x, b, act=act, alpha=alpha, gain=gain, clamp=clamp)
        np.testing.assert_allclose(y.detach().cpu().numpy(), g[f'case{i}_y'], rtol=2e-6, atol=2e-6, err_msg=f'fwd {act}')
        dy = t(g[f'case{i}_dy']).to(DEV)
        dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
        np.testing.assert_allclose(dx.detach().cpu().numpy(), g[f'case{i}_dx'], rtol=2e-5, atol=2e-5, err_msg=f'dx {act}')
        np.testing.assert_allclose(db.detach().cpu().numpy(), g[f'case{i}_db'], rtol=1e-4, atol=1e-4, err_msg=f'db {act}')
        if bias_act.activation_funcs[act].has_2nd_grad:
            v = t(g[f'case{i}_v']).to(DEV)
            ddx, = torch.autograd.grad(dx, [x], v)
            np.testing.assert_allclose(ddx.cpu().numpy(), g[f'case{i}_ddx'], rtol=1e-4, atol=1e-4, err_msg=f'ddx {act}')
    assert custom_ops.launch_count() > before        # the native library did the work
    x2, b2 = t(g['x2']).to(DEV), t(g['b2']).to(DEV)
    np.testing.assert_allclose(bias_act.bias_act(x2, b2, dim=1, act='lrelu').cpu().numpy(), g['y2'], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(bias_act.bias_act(x2, None, act='relu', gain=1.0).cpu().numpy(), g['y2_nob'], rtol=0, atol=0)
    np.testing.assert_allclose(bias_act.bias_act(t(g['x']).to(DEV), t(g['b3']).to(DEV), dim=3, act='swish').cpu().numpy(),
                               g['y3_lastdim'], rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.float64, 1e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize('act', list(ref_ops.ACTIVATIONS))
def test_all_activations_all_dtypes_vs_oracle(act, dtype, tol):
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(3, 6, 17, 19, generator=g, dtype=torch.float64) * 2).to(dtype)
    b = torch.randn(6, generator=g, dtype=torch.float64).to(dtype)
    want = ref_ops.bias_act(x.double(), b.double(), act=act, clamp=3.0)
    got = bias_act.bias_act(x.to(DEV), b.to(DEV), act=act, clamp=3.0)
    assert got.dtype == dtype and got.shape == x.shape
    assert max_abs(got, want) <= tol * 4, (act, dtype, max_abs(got, want))


@pytest.mark.parametrize('shape,dim', [((4, 512), 1), ((5, 7, 3), 0), ((2, 9, 5, 5), 2), ((1, 3, 2, 1031), 3), ((7,), 0), ((2, 8, 33, 35), 1)])
def test_bias_dims_tails_and_unaligned_views(shape, dim):
    g = torch.Generator().manual_seed(12)
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[dim], generator=g)
    want = ref_ops.bias_act(x, b, dim=dim, act='lrelu', gain=1.3)
    got = bias_act.bias_act(x.to(DEV), b.to(DEV), dim=dim, act='lrelu', gain=1.3)
    assert max_abs(got, want) < 3e-6
    # storage offset that breaks 16-byte alignment -> scalar kernel
    buf = torch.zeros(x.numel() + 1, device=DEV)
    xv = buf[1:].view(*shape); xv.copy_(x)
    assert max_abs(bias_act.bias_act(xv, b.to(DEV), dim=dim, act='lrelu', gain=1.3), want) < 3e-6


def test_channels_last_and_empty_and_identity():
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 12, 9, 7, generator=g)
    b = torch.randn(12, generator=g)
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    y = bias_act.bias_act(xc, b.to(DEV), act='relu')
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert max_abs(y, ref_ops.bias_act(x, b, act='relu')) < 3e-6
    e = torch.empty(0, 4, 3, 3, device=DEV)
    assert bias_act.bias_act(e, torch.zeros(4, device=DEV), act='lrelu').shape == e.shape
    xi = x.to(DEV)
    assert bias_act.bias_act(xi, None, act='linear', gain=1).data_ptr() == xi.data_ptr()      # identity returns x (bias_act.py:151-153)


def test_argument_errors_raise():
    x = torch.randn(2, 3, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        bias_act.bias_act(x, torch.zeros(5, device=DEV))                    # wrong number of elements
    with pytest.raises(RuntimeError):
        bias_act.bias_act(x, torch.zeros(3, device=DEV, dtype=torch.float16))   # dtype mismatch
    with pytest.raises(RuntimeError):
        bias_act.bias_act(x, torch.zeros(3, device=DEV), dim=7)
    with pytest.raises(KeyError):
        bias_act.bias_act(x, None, act='gelu')


def test_full_size_generator_tensor_properties():
    # [8, 64, 512, 512] fp32 (the generator's biggest activation, a quarter of the batch-32 config):
    # checked against torch on the same device, plus clamp-range and sign properties
    torch.manual_seed(0)
    x = torch.randn(8, 64, 512, 512, device=DEV) * 100
    b = torch.randn(64, device=DEV)
    y = bias_act.bias_act(x, b, act='lrelu', gain=float(np.sqrt(2)), clamp=256.0)
    want = (torch.nn.functional.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * float(np.sqrt(2))).clamp(-256, 256)
    assert float((y - want).abs().max()) < 1e-4
    assert float(y.max()) <= 256.0 and float(y.min()) >= -256.0
    assert bool(((y > 0) == ((x + b.view(1, -1, 1, 1)) > 0)).all())
