"""`modulated_conv2d` -- the StyleGAN2 modulated convolution of PASTA-GAN++ -- on the B200 tensor cores.

Drop-in for the reference's training/networks.py:36-94 (same signature and argument meaning).
The reference either materialises per-sample weights and runs a `groups=batch` convolution
(fused_modconv=True) or scales activations before and after a shared-weight convolution
(fused_modconv=False); both compute

    y[n,o] = d[n,o] * sum_{i,k} w[o,i,k] * s[n,i] * x[n,i] (+ noise),   d = rsqrt(sum_{i,k} (w*s)^2 + 1e-8)

On CUDA tensors this module evaluates that expression as ONE shared-weight implicit GEMM over all
samples: the style scale s is folded into the activation packing pass, the demodulation scale d and the
noise into the GEMM epilogue, and the up=2 resampling into polyphase weights.  No per-sample weight
tensor is ever written.  `fused_modconv` therefore only selects the formulation on the differentiable /
non-CUDA composition path, as in the reference.
"""
import numpy as np
import torch

from ..torch_utils import misc
from ..torch_utils.ops import bias_act
from ..torch_utils.ops import conv2d_gradfix
from ..torch_utils.ops import conv2d_resample
from ..torch_utils.ops import fma
from ..torch_utils.ops import upfirdn2d


def _kernel_path_ok(x, weight, styles, noise, up, down, padding, resample_filter):
    if not conv2d_gradfix._should_use_custom_op(x):
        return False
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, weight, styles, noise)):
        return False
    kh, kw = weight.shape[2:]
    if down != 1 or kh != kw:
        return False
    if up == 1:
        return isinstance(padding, int) and 0 <= padding <= kh - 1
    if up == 2:
        f = resample_filter
        return kh == 3 and padding == 1 and f is not None and f.ndim == 2 and tuple(f.shape) == (4, 4)
    return False


def modulated_conv2d_fused_act(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None, demodulate=True,
                               flip_weight=True, bias=None, act='linear', alpha=None, gain=None, clamp=None,
                               out=None, out_dtype=None, accumulate=False, memory_format=None, out_packed=None):
    """Kernel path of modulated_conv2d with bias_act fused into the same launch (what SynthesisLayer / ToRGB
    compose, networks.py:1925-1935 and SURVEY Appendix E): returns
    clamp(act(modconv(x) + bias) * gain).  Inference only (no autograd).
    x may be a conv2d_gradfix.PackedAct (then the style scale is folded into per-sample weights instead of the packing
    pass) and out_packed a PackedAct view to receive the result in operand format."""
    spec = bias_act.activation_funcs[act]
    alpha = float(spec.def_alpha if alpha is None else alpha)
    gain = float(spec.def_gain if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    src_dtype = torch.float32 if isinstance(x, conv2d_gradfix.PackedAct) else x.dtype
    _, parts = conv2d_gradfix._PRODUCTS[conv2d_gradfix.precision_for(src_dtype)]
    dcoef = None
    if demodulate:
        conv2d_gradfix._init()
        dcoef = conv2d_gradfix._plugin.demod_coefs(weight, styles)
    if up == 2:
        pw = conv2d_gradfix.packed_up2(weight, resample_filter, flip_weight, False, parts)
    else:
        pw = conv2d_gradfix.packed_plain(weight, flip_weight, parts, padding, padding)
    return conv2d_gradfix.igemm_conv(x, pw, scale=styles, dcoef=dcoef, noise=noise, bias=bias, act=act, alpha=alpha,
                                     gain=gain, clamp=clamp, out=out, out_dtype=out_dtype, accumulate=accumulate,
                                     memory_format=memory_format, out_packed=out_packed)


@misc.profiled_function
def modulated_conv2d(
    x,                          # Input tensor of shape [batch_size, in_channels, in_height, in_width].
    weight,                     # Weight tensor of shape [out_channels, in_channels, kernel_height, kernel_width].
    styles,                     # Modulation coefficients of shape [batch_size, in_channels].
    noise           = None,     # Optional noise tensor to add to the output activations.
    up              = 1,        # Integer upsampling factor.
    down            = 1,        # Integer downsampling factor.
    padding         = 0,        # Padding with respect to the upsampled image.
    resample_filter = None,     # Low-pass filter from upfirdn2d.setup_filter().
    demodulate      = True,     # Apply weight demodulation?
    flip_weight     = True,     # False = convolution, True = correlation (matches torch.nn.functional.conv2d).
    fused_modconv   = True,     # Formulation used on the composition path (see module docstring).
):
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    misc.assert_shape(weight, [out_channels, in_channels, kh, kw])
    misc.assert_shape(x, [batch_size, in_channels, None, None])
    misc.assert_shape(styles, [batch_size, in_channels])

    # ---- sm_100a kernel path: one implicit GEMM, modulation / demodulation / noise fused ----
    if _kernel_path_ok(x, weight, styles, noise, up, down, padding, resample_filter):
        return modulated_conv2d_fused_act(x, weight, styles, noise=noise, up=up, padding=padding,
                                          resample_filter=resample_filter, demodulate=demodulate, flip_weight=flip_weight)

    # ---- composition path (differentiable; any device): the reference's two formulations ----
    if x.dtype == torch.float16 and demodulate:     # pre-normalise to avoid fp16 overflow
        weight = weight * (1 / np.sqrt(in_channels * kh * kw) / weight.norm(float('inf'), dim=[1, 2, 3], keepdim=True))
        styles = styles / styles.norm(float('inf'), dim=1, keepdim=True)
    w = dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(batch_size, 1, -1, 1, 1)           # [N, O, I, kh, kw]
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()                     # [N, O]
    use_grouped = fused_modconv and not conv2d_gradfix._should_use_custom_op(x)     # grouped conv only via the library
    if not use_grouped:
        x = x * styles.to(x.dtype).reshape(batch_size, -1, 1, 1)
        x = conv2d_resample.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down,
                                            padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma.fma(x, dcoefs.to(x.dtype).reshape(batch_size, -1, 1, 1), noise.to(x.dtype))
        elif demodulate:
            x = x * dcoefs.to(x.dtype).reshape(batch_size, -1, 1, 1)
        elif noise is not None:
            x = x.add_(noise.to(x.dtype))
        return x
    if demodulate:
        w = w * dcoefs.reshape(batch_size, -1, 1, 1, 1)
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, in_channels, kh, kw)
    x = conv2d_resample.conv2d_resample(x=x, w=w.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                                        groups=batch_size, flip_weight=flip_weight)
    x = x.reshape(batch_size, -1, *x.shape[2:])
    if noise is not None:
        x = x.add_(noise)
    return x
