"""Fused bias + activation + gain + clamp.  Drop-in for the reference's torch_utils/ops/bias_act.py:
same `bias_act(x, b, dim, act, alpha, gain, clamp, impl)` signature, same `activation_funcs` table,
same first/second-order gradient structure (bias_act.py:129-210), served by the sm_100a kernel in
csrc/bias_act.cu through the C ABI (`pgpp_bias_act`).

Differences from the reference, on purpose:
  * impl='cuda' never falls back: a non-CUDA tensor or a missing library raises RuntimeError
    (the reference silently takes the ref path, bias_act.py:87).  impl='ref' is plain PyTorch on any device.
  * bfloat16 is supported in addition to float16/32/64.
  * for act='linear' with a clamp the saved output gates the gradient, as the ref path's autograd does
    (the reference's CUDA path passes an empty yref there, bias_act.py:154-157, and lets it through).
"""
import numpy as np
import torch

from .. import custom_ops
from .. import misc


class _Spec(dict):
    """attribute-style record (stands in for dnnlib.EasyDict)"""
    __getattr__ = dict.__getitem__


def _spec(func, def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
    return _Spec(func=func, def_alpha=def_alpha, def_gain=def_gain, cuda_idx=cuda_idx, ref=ref, has_2nd_grad=has_2nd_grad)


_F = torch.nn.functional
activation_funcs = {    # bias_act.py:23-33
    'linear':   _spec(lambda x, **_: x,                            0,   1,          1, '',  False),
    'relu':     _spec(lambda x, **_: _F.relu(x),                   0,   np.sqrt(2), 2, 'y', False),
    'lrelu':    _spec(lambda x, alpha, **_: _F.leaky_relu(x, alpha), 0.2, np.sqrt(2), 3, 'y', False),
    'tanh':     _spec(lambda x, **_: torch.tanh(x),                0,   1,          4, 'y', True),
    'sigmoid':  _spec(lambda x, **_: torch.sigmoid(x),             0,   1,          5, 'y', True),
    'elu':      _spec(lambda x, **_: _F.elu(x),                    0,   1,          6, 'y', True),
    'selu':     _spec(lambda x, **_: _F.selu(x),                   0,   1,          7, 'y', True),
    'softplus': _spec(lambda x, **_: _F.softplus(x),               0,   1,          8, 'y', True),
    'swish':    _spec(lambda x, **_: torch.sigmoid(x) * x,         0,   np.sqrt(2), 9, 'x', True),
}

_plugin = None
_null_tensor = torch.empty([0])


def _init():
    """Load the prebuilt plugin (raises if unavailable; no fallback)."""
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('bias_act_plugin')
    return True


def _resolve(act, alpha, gain, clamp):
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    return (spec, float(spec.def_alpha if alpha is None else alpha), float(spec.def_gain if gain is None else gain),
            float(-1 if clamp is None else clamp))


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """y = clamp(act(x + b) * gain, -clamp, clamp); supports first and second order gradients."""
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl == 'ref':
        return _bias_act_ref(x=x, b=b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
    if x.device.type != 'cuda':
        if custom_ops.cpu_tensors == 'ref':     # explicit opt-in to the reference's dispatch rule (bias_act.py:87)
            return _bias_act_ref(x=x, b=b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
        raise RuntimeError("bias_act(impl='cuda') needs a CUDA tensor; pass impl='ref' for the PyTorch reference path")
    _init()
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


@misc.profiled_function
def _bias_act_ref(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Plain PyTorch ops on any device (bias_act.py:93-123)."""
    assert isinstance(x, torch.Tensor)
    spec, alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
        view = [1] * x.ndim
        view[dim] = -1
        x = x + b.reshape(view)
    y = spec.func(x, alpha=alpha)
    if gain != 1:
        y = y * gain
    if clamp >= 0:
        y = y.clamp(-clamp, clamp)
    return y


_bias_act_cuda_cache = dict()


def _memory_format(t):
    return torch.channels_last if t.ndim > 2 and t.stride()[1] == 1 else torch.contiguous_format


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Builds (and caches) the autograd.Function pair for one parameter set (bias_act.py:129-210)."""
    spec, alpha, gain, clamp = _resolve(act, alpha, gain, clamp)
    key = (dim, act, alpha, gain, clamp)
    if key in _bias_act_cuda_cache:
        return _bias_act_cuda_cache[key]
    idx = spec.cuda_idx
    is_identity = (act == 'linear' and gain == 1 and clamp < 0)
    keep_x = 'x' in spec.ref or spec.has_2nd_grad
    keep_y = 'y' in spec.ref or (act == 'linear' and clamp >= 0)
    other_dims = lambda t: [i for i in range(t.ndim) if i != dim]

    def bias_grad(t):
        """t.sum(over every dimension but `dim`) (bias_act.py:135,156); for NCHW-contiguous CUDA images one pass of pgpp_sum_hw
        (float32 accumulation) + a [N, C] column sum instead of the library's generic reduction"""
        if (t.is_cuda and t.ndim == 4 and dim == 1 and t.dtype in (torch.float32, torch.float16, torch.bfloat16) and t.is_contiguous()
                and t.shape[2] * t.shape[3] >= 1024 and not (torch.is_grad_enabled() and t.requires_grad)):
            return custom_ops.get_plugin('conv2d_plugin').sum_hw(t).sum(0).to(t.dtype)
        return t.sum(other_dims(t))

    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b):
            ctx.memory_format = _memory_format(x)
            x = x.contiguous(memory_format=ctx.memory_format)
            b = b.contiguous() if b is not None else _null_tensor
            if is_identity and b is _null_tensor:
                y = x
            else:
                y = _plugin.bias_act(x, b, _null_tensor, _null_tensor, _null_tensor, 0, dim, idx, alpha, gain, clamp)
            ctx.save_for_backward(x if keep_x else _null_tensor, b if keep_x else _null_tensor, y if keep_y else _null_tensor)
            return y

        @staticmethod
        def backward(ctx, dy):
            dy = dy.contiguous(memory_format=ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = db = None
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = dy if is_identity else BiasActCudaGrad.apply(dy, x, b, y)
            if ctx.needs_input_grad[1]:
                db = bias_grad(dx)
            return dx, db

    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y):
            ctx.memory_format = _memory_format(dy)
            dx = _plugin.bias_act(dy, b, x, y, _null_tensor, 1, dim, idx, alpha, gain, clamp)
            ctx.save_for_backward(dy if spec.has_2nd_grad else _null_tensor, x, b, y)
            return dx

        @staticmethod
        def backward(ctx, d_dx):
            d_dx = d_dx.contiguous(memory_format=ctx.memory_format)
            dy, x, b, y = ctx.saved_tensors
            d_dy = d_x = d_b = None
            if ctx.needs_input_grad[0]:
                d_dy = BiasActCudaGrad.apply(d_dx, x, b, y)
            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                d_x = _plugin.bias_act(d_dx, b, x, y, dy, 2, dim, idx, alpha, gain, clamp)
            if spec.has_2nd_grad and ctx.needs_input_grad[2]:
                d_b = bias_grad(d_x)
            return d_dy, d_x, d_b, None

    BiasActCuda.Grad, BiasActCuda.is_identity, BiasActCuda.keep_y = BiasActCudaGrad, is_identity, keep_y
    _bias_act_cuda_cache[key] = BiasActCuda
    return BiasActCuda
