"""Pad / upsample / FIR-filter / downsample 2-D images.  Drop-in for the reference's
torch_utils/ops/upfirdn2d.py (same public functions and signatures), served by csrc/upfirdn2d.cu
through `pgpp_upfirdn2d`.

impl='cuda' never falls back (non-CUDA tensor or missing library -> RuntimeError); impl='ref' is a
plain-PyTorch path usable on any device.  bfloat16 is accepted in addition to float16/32/64.
"""
import math

import torch

from .. import custom_ops
from .. import misc

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('upfirdn2d_plugin')
    return True


# ---- argument forms of the public API (upfirdn2d.py:22-56 of the reference: an int, or a list / tuple of ints) ------------------

def _int_list(value, what):
    if isinstance(value, int):
        return [value, value]
    assert isinstance(value, (list, tuple)), f'{what} must be an int or a list / tuple of ints'
    assert all(isinstance(v, int) for v in value), f'{what} must hold ints'
    return list(value)


def _parse_scaling(scaling):
    """int or [x, y] -> (sx, sy), both >= 1"""
    sx, sy = _int_list(scaling, 'scaling')
    assert min(sx, sy) >= 1
    return sx, sy


def _parse_padding(padding):
    """int, [x, y] or [x_before, x_after, y_before, y_after] -> the four-element form"""
    values = _int_list(padding, 'padding')
    if len(values) == 2:
        values = [values[0], values[0], values[1], values[1]]
    x_before, x_after, y_before, y_after = values
    return x_before, x_after, y_before, y_after


def _get_filter_size(f):
    """(width, height) of a filter tensor; (1, 1) for None = identity"""
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2
    height, width = int(f.shape[0]), int(f.shape[-1])
    assert min(width, height) >= 1
    return width, height


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """Prepare a float32 FIR filter: [fh, fw] (non-separable) or [taps] (separable; automatic for 1-D with
    at least 8 taps).  See the reference's upfirdn2d.py:72-116 for the contract."""
    taps = torch.as_tensor(1 if f is None else f, dtype=torch.float32)
    if taps.ndim == 0:
        taps = taps.reshape(1)
    assert taps.ndim in (1, 2) and taps.numel() > 0
    if separable is None:
        separable = taps.ndim == 1 and taps.numel() >= 8
    if taps.ndim == 1 and not separable:
        taps = torch.outer(taps, taps)
    assert taps.ndim == (1 if separable else 2)
    if normalize:
        taps = taps / taps.sum()
    if flip_filter:
        taps = taps.flip(list(range(taps.ndim)))
    taps = taps * (gain ** (taps.ndim / 2))        # a separable filter is applied twice: sqrt(gain) per pass
    return taps.to(device=device)


# ---- the op -----------------------------------------------------------------------------------------------------------------------

def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Upsample by zero insertion (`up`), pad / crop (`padding`), convolve with `f`, keep every `down`-th
    pixel.  x: [N, C, H, W]; f: float32 [fh, fw], [taps] (separable) or None (identity)."""
    assert isinstance(x, torch.Tensor) and impl in ('ref', 'cuda')
    args = dict(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain)
    if impl == 'cuda' and x.device.type == 'cuda':
        _init()
        return _upfirdn2d_cuda(**args).apply(x, f)
    if impl == 'cuda' and custom_ops.cpu_tensors != 'ref':     # 'ref': explicit opt-in to the reference's dispatch rule (upfirdn2d.py:162)
        raise RuntimeError("upfirdn2d(impl='cuda') needs a CUDA tensor; pass impl='ref' for the PyTorch reference path")
    return _upfirdn2d_ref(x, f, **args)


@misc.profiled_function
def _upfirdn2d_ref(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Plain PyTorch ops (zero insertion, F.pad, depthwise conv2d, slicing); any device."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    if f is None:
        f = x.new_ones([1, 1], dtype=torch.float32)
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2) and f.dtype == torch.float32 and not f.requires_grad
    n, c, h, w = x.shape
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    pad = torch.nn.functional.pad

    x = pad(x.reshape(n, c, h, 1, w, 1), [0, upx - 1, 0, 0, 0, upy - 1]).reshape(n, c, h * upy, w * upx)
    x = pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]

    k = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:         # conv2d correlates: a true convolution needs the flipped taps
        k = k.flip(list(range(k.ndim)))
    conv = torch.nn.functional.conv2d
    if k.ndim == 2:
        x = conv(x, k[None, None].repeat(c, 1, 1, 1), groups=c)
    else:
        x = conv(x, k[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
        x = conv(x, k[None, None, :, None].repeat(c, 1, 1, 1), groups=c)
    return x[:, :, ::downy, ::downx]


_upfirdn2d_cuda_cache = dict()


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    """autograd.Function for one parameter set (cached); the backward pass is the same op with up / down swapped,
    the filter flipped and the padding of upfirdn2d.py:251-256."""
    ups, downs, pads = _parse_scaling(up), _parse_scaling(down), _parse_padding(padding)
    key = ups + downs + pads + (flip_filter, gain)
    fn = _upfirdn2d_cuda_cache.get(key)
    if fn is None:
        fn = _upfirdn2d_cuda_cache[key] = _make_function(ups, downs, pads, flip_filter, gain)
    return fn


def _run_kernels(x, f, ups, downs, pads, flip_filter, gain):
    """one launch for a 2-D filter; for a separable one a horizontal and a vertical pass with sqrt(gain) each"""
    (upx, upy), (downx, downy), (px0, px1, py0, py1) = ups, downs, pads
    if f.ndim == 2:
        return _plugin.upfirdn2d(x, f, upx, upy, downx, downy, px0, px1, py0, py1, flip_filter, gain)
    g = math.sqrt(gain)
    y = _plugin.upfirdn2d(x, f.unsqueeze(0), upx, 1, downx, 1, px0, px1, 0, 0, flip_filter, g)
    return _plugin.upfirdn2d(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, py0, py1, flip_filter, g)


def _make_function(ups, downs, pads, flip_filter, gain):
    (upx, upy), (downx, downy), (px0, _, py0, _) = ups, downs, pads

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = x.new_ones([1, 1], dtype=torch.float32)
            assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
            ctx.save_for_backward(f)
            ctx.in_hw = tuple(x.shape[2:])
            return _run_kernels(x, f, ups, downs, pads, flip_filter, gain)

        @staticmethod
        def backward(ctx, dy):
            assert not ctx.needs_input_grad[1], 'the filter is a constant'
            if not ctx.needs_input_grad[0]:
                return None, None
            f, = ctx.saved_tensors
            (ih, iw), (oh, ow) = ctx.in_hw, dy.shape[2:]
            fw, fh = _get_filter_size(f)
            # the adjoint: up and down trade places, the filter is flipped, and the padding restores the input size (upfirdn2d.py:251-256)
            adjoint_pad = [fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1,
                           fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1]
            adjoint = _upfirdn2d_cuda(up=list(downs), down=list(ups), padding=adjoint_pad, flip_filter=(not flip_filter), gain=gain)
            return adjoint.apply(dy, f), None

    return Upfirdn2dCuda


# ---- wrappers with "same"-size semantics (upfirdn2d.py:264-384) -------------------------------------------------------------------

def _widen(padding, f, before, after):
    """user padding plus the filter's own support: `before(taps, axis)` / `after(taps, axis)` pixels in front of / behind each axis"""
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    return [px0 + before(fw, 0), px1 + after(fw, 0), py0 + before(fh, 1), py1 + after(fh, 1)]


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """FIR-filter with "same" output size (user padding on top)."""
    p = _widen(padding, f, lambda t, _: t // 2, lambda t, _: (t - 1) // 2)
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Upsample so that the output is `up` times the input size (user padding on top)."""
    factor = _parse_scaling(up)
    p = _widen(padding, f, lambda t, ax: (t + factor[ax] - 1) // 2, lambda t, ax: (t - factor[ax]) // 2)
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * factor[0] * factor[1], impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Downsample so that the output is 1/`down` of the input size (user padding on top)."""
    factor = _parse_scaling(down)
    p = _widen(padding, f, lambda t, ax: (t - factor[ax] + 1) // 2, lambda t, ax: (t - factor[ax]) // 2)
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
