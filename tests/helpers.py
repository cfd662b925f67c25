"""Shared test helpers (CPU emulation of the packed-operand GEMM, tolerances)."""
import numpy as np
import torch
import torch.nn.functional as F


def t(a):
    return torch.from_numpy(np.asarray(a))


def rel_l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def emulate_igemm(x, pw, scale=None, stride=1, parts=None):
    """What csrc/conv_igemm.cu computes from packed weights, in float64 on the CPU:
    out[n, g, y, x] = sum_{tap,c} xs[n, c, y*s+ky-pad_y, x*s+kx-pad_x] * W[tap][g][c], then the phase scatter."""
    parts = parts or pw.parts
    w = pw.data[:parts].double().sum(0).cpu()                # [taps, o_rows, c_pad]
    n, c, h, wd = x.shape
    xs = x.double().cpu()
    if scale is not None:
        xs = xs * scale.double().cpu().reshape(n, c, 1, 1)
    cols = pw.phases * pw.o
    wsel = w[:, :pw.phases * pw.phase_stride].reshape(w.shape[0], pw.phases, pw.phase_stride, -1)[:, :, :pw.o, :c]
    wk = wsel.reshape(pw.kh, pw.kw, cols, c).permute(2, 3, 0, 1)      # [cols, c, kh, kw]
    xp = F.pad(xs, [pw.pad_x, pw.kw, pw.pad_y, pw.kh])      # generous bottom/right zero padding
    y = F.conv2d(xp, wk, stride=stride)
    conv_h = (h + 2 * pw.pad_y - pw.kh) // stride + 1
    conv_w = (wd + 2 * pw.pad_x - pw.kw) // stride + 1
    y = y[:, :, :conv_h, :conv_w]
    if pw.phases == 4:
        y = y.reshape(n, 2, 2, pw.o, conv_h, conv_w).permute(0, 3, 4, 1, 5, 2).reshape(n, pw.o, conv_h * 2, conv_w * 2)
    return y


import contextlib


@contextlib.contextmanager
def upfirdn2d_ref_on_cpu(upfirdn2d_module):
    """CPU tests of the callers: conv2d_resample reaches upfirdn2d with the default impl='cuda', which by design
    refuses CPU tensors; route those calls to the module's own PyTorch path for the duration of the test."""
    orig = upfirdn2d_module.upfirdn2d
    upfirdn2d_module.upfirdn2d = lambda *a, **k: orig(*a, **{**k, 'impl': 'ref'})
    try:
        yield
    finally:
        upfirdn2d_module.upfirdn2d = orig
