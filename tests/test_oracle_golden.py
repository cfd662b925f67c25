"""Pins oracle/ref_ops.py (the CPU restatement) to fixtures minted from the real reference
(oracle/make_golden.py) and to the plain-numpy definition of the third-party convolutions."""
import os

import numpy as np
import pytest
import torch

from oracle import np_kernels, ref_ops
from oracle.make_golden import CONV_CASES, MODCONV_CASES, UPFIRDN_CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
torch.set_num_threads(1)


def _load(name):
    return np.load(os.path.join(GOLDEN, name))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _opt(v):
    return None if v == 'None' else float(v)


def test_bias_act_forward_and_grads_match_reference():
    g = _load('bias_act.npz')
    x, b = _t(g['x']), _t(g['b'])
    ncases = len([k for k in g.files if k.endswith('_meta')])
    assert ncases == 22
    for i in range(ncases):
        act, alpha, gain, clamp = g[f'case{i}_meta']
        alpha, gain, clamp = _opt(alpha), _opt(gain), _opt(clamp)
        y = ref_ops.bias_act(x, b, dim=1, act=act, alpha=alpha, gain=gain, clamp=clamp)
        np.testing.assert_allclose(y.numpy(), g[f'case{i}_y'], rtol=1e-6, atol=1e-6, err_msg=f'fwd {act}')
        # plugin-style arithmetic: grad=0 equals forward; grad=1 equals autograd of the ref path
        a, gn, cl, *_ = ref_ops.resolve_act(act, alpha, gain, clamp)
        y0 = ref_ops.bias_act_plugin(x, b, None, None, None, 0, 1, act, a, gn, cl)
        np.testing.assert_allclose(y0.numpy(), g[f'case{i}_y'], rtol=2e-6, atol=2e-6, err_msg=f'plugin fwd {act}')
        dy = _t(g[f'case{i}_dy'])
        yref = _t(g[f'case{i}_y'])
        dx = ref_ops.bias_act_plugin(dy, b, x, yref, None, 1, 1, act, a, gn, cl)
        np.testing.assert_allclose(dx.numpy(), g[f'case{i}_dx'], rtol=2e-5, atol=2e-5, err_msg=f'grad1 {act}')
        np.testing.assert_allclose(dx.sum(dim=(0, 2, 3)).numpy(), g[f'case{i}_db'], rtol=1e-4, atol=1e-4)
        if ref_ops.ACTIVATIONS[act][4]:
            v = _t(g[f'case{i}_v'])
            ddx = ref_ops.bias_act_plugin(v, b, x, yref, dy, 2, 1, act, a, gn, cl)
            np.testing.assert_allclose(ddx.numpy(), g[f'case{i}_ddx'], rtol=5e-5, atol=5e-5, err_msg=f'grad2 {act}')


def test_bias_act_dims_match_reference():
    g = _load('bias_act.npz')
    x2, b2 = _t(g['x2']), _t(g['b2'])
    np.testing.assert_allclose(ref_ops.bias_act(x2, b2, dim=1, act='lrelu').numpy(), g['y2'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ref_ops.bias_act(x2, None, act='relu', gain=1.0).numpy(), g['y2_nob'], rtol=0, atol=0)
    np.testing.assert_allclose(ref_ops.bias_act(_t(g['x']), _t(g['b3']), dim=3, act='swish').numpy(), g['y3_lastdim'],
                               rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('case', UPFIRDN_CASES, ids=[c[0] for c in UPFIRDN_CASES])
def test_upfirdn2d_matches_reference(case):
    name, shape, taps, sep, up, down, pad, flip, gain = case
    g = _load('upfirdn2d.npz')
    x = _t(g[f'{name}_x'])
    f = _t(g[f'{name}_f']) if g[f'{name}_f'].size else None
    y = ref_ops.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
    assert tuple(y.shape) == g[f'{name}_y'].shape
    fw, fh = ref_ops.filter_size(f)
    assert tuple(y.shape[2:]) == ref_ops.upfirdn2d_out_size(x.shape[2], x.shape[3], fh, fw, up, down, pad)
    np.testing.assert_allclose(y.numpy(), g[f'{name}_y'], rtol=1e-5, atol=2e-6)
    # the gradient is the same operator with swapped factors (upfirdn2d.py:246-261)
    dy = _t(g[f'{name}_dy'])
    kw = ref_ops.upfirdn2d_backward_args(x.shape, y.shape, f, up, down, pad, flip, gain)
    dx = ref_ops.upfirdn2d(dy, f, **kw)
    np.testing.assert_allclose(dx.numpy(), g[f'{name}_dx'], rtol=1e-5, atol=2e-6)


def test_upfirdn2d_wrappers_and_setup_filter():
    g = _load('upfirdn2d.npz')
    x, f = _t(g['wrap_x']), _t(g['wrap_f'])
    np.testing.assert_allclose(ref_ops.filter2d(x, f).numpy(), g['wrap_filter2d'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ref_ops.upsample2d(x, f).numpy(), g['wrap_upsample2d'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ref_ops.downsample2d(x, f).numpy(), g['wrap_downsample2d'], rtol=1e-5, atol=1e-6)
    from oracle.make_golden import SYM6
    np.testing.assert_allclose(ref_ops.setup_filter([1, 3, 3, 1]).numpy(), g['setup_1331'], rtol=1e-7)
    np.testing.assert_allclose(ref_ops.setup_filter(SYM6).numpy(), g['setup_sym6'], rtol=1e-7)
    np.testing.assert_allclose(ref_ops.setup_filter([1, 2, 3], flip_filter=True, gain=3.0).numpy(), g['setup_flip_gain'], rtol=1e-6)


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv2d_resample_matches_reference(case):
    name, xs, ws, up, down, pad, groups, flipw, usef = case
    g = _load('conv2d_resample.npz')
    x, w, f = _t(g[f'{name}_x']), _t(g[f'{name}_w']), _t(g['f'])
    y = ref_ops.conv2d_resample(x, w, f=(f if usef else None), up=up, down=down, padding=pad, groups=groups,
                                flip_weight=flipw)
    np.testing.assert_allclose(y.numpy(), g[f'{name}_y'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('case', MODCONV_CASES, ids=[c[0] for c in MODCONV_CASES])
def test_modulated_conv2d_matches_reference(case):
    name, n, ic, oc, k, h, up, demod, noise_kind, flipw = case
    g = _load('modulated_conv2d.npz')
    x, w, s, f = _t(g[f'{name}_x']), _t(g[f'{name}_w']), _t(g[f'{name}_s']), _t(g['f'])
    noise = _t(g[f'{name}_noise']) if g[f'{name}_noise'].size else None
    kw = dict(noise=noise, up=up, padding=k // 2, resample_filter=f, demodulate=demod, flip_weight=flipw)
    for fused, key in ((True, 'y_fused'), (False, 'y_split')):
        y = ref_ops.modulated_conv2d(x, w, s, fused_modconv=fused, **kw)
        ref = g[f'{name}_{key}']
        err = np.abs(y.numpy() - ref).max() / np.abs(ref).max()
        assert err < 2e-6, (name, fused, err)
    b = _t(g[f'{name}_b'])
    if demod:
        ya = ref_ops.synthesis_layer(x, s, w, b, noise, up, f)
    else:
        ya = ref_ops.to_rgb(x, s, w, b)
    np.testing.assert_allclose(ya.numpy(), g[f'{name}_y_act'], rtol=1e-4, atol=1e-4)


def test_library_convolutions_match_plain_numpy_definition():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 7, 8, generator=g, dtype=torch.float64)
    w = torch.randn(4, 3, 3, 3, generator=g, dtype=torch.float64)
    for stride, pad in ((1, 1), (2, 0), (2, 1)):
        a = ref_ops.conv2d(x, w, stride=stride, padding=pad, groups=2).numpy()
        b = np_kernels.conv2d(x.numpy(), w.numpy(), stride=stride, padding=(pad, pad), groups=2)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    wt = torch.randn(6, 2, 3, 3, generator=g, dtype=torch.float64)
    for stride, pad in ((1, 0), (2, 0), (2, 1)):
        a = ref_ops.conv_transpose2d(x, wt, stride=stride, padding=pad, groups=2).numpy()
        b = np_kernels.conv_transpose2d(x.numpy(), wt.numpy(), stride=stride, padding=(pad, pad), groups=2)
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
