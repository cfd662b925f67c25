"""Plain-numpy definitions of the third-party convolution arithmetic (TEST INFRASTRUCTURE ONLY).

The reference's convolutions are calls into PyTorch (`torch.nn.functional.conv2d`,
`conv_transpose2d`; call sites torch_utils/ops/conv2d_gradfix.py:38,43,112,114 and
torch_utils/ops/upfirdn2d.py:201-204).  PyTorch is not part of /root/reference, so this file
restates the published definition of those two operators with explicit loops over the filter
taps, in float64, for small cases.  `tests/test_oracle_golden.py` pins `oracle/ref_ops.py`
(which calls the library) against these loops.
"""
import numpy as np


def conv2d(x, w, stride=1, padding=(0, 0), groups=1):
    """out[n,o,y,x] = sum_{i,ky,kx} in[n, g*Ig+i, y*s+ky-py, x*s+kx-px] * w[o,i,ky,kx]  (cross-correlation)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, c, h, wd = x.shape
    oc, icg, kh, kw = w.shape
    py, px = padding
    assert c == icg * groups and oc % groups == 0
    xp = np.zeros((n, c, h + 2 * py, wd + 2 * px))
    xp[:, :, py:py + h, px:px + wd] = x
    oh = (h + 2 * py - kh) // stride + 1
    ow = (wd + 2 * px - kw) // stride + 1
    out = np.zeros((n, oc, oh, ow))
    ocg = oc // groups
    for g in range(groups):
        xs = xp[:, g * icg:(g + 1) * icg]
        ws = w[g * ocg:(g + 1) * ocg]
        for ky in range(kh):
            for kx in range(kw):
                patch = xs[:, :, ky:ky + (oh - 1) * stride + 1:stride, kx:kx + (ow - 1) * stride + 1:stride]
                out[:, g * ocg:(g + 1) * ocg] += np.einsum('nihw,oi->nohw', patch, ws[:, :, ky, kx])
    return out


def conv_transpose2d(x, w, stride=1, padding=(0, 0), groups=1):
    """Adjoint of conv2d: every input pixel scatters w[i,o,:,:] into out[y*s+ky-py, x*s+kx-px].
    w has shape [in_channels, out_channels/groups, kh, kw]."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    n, c, h, wd = x.shape
    ic, ocg, kh, kw = w.shape
    py, px = padding
    assert ic == c and c % groups == 0
    icg = c // groups
    fh = (h - 1) * stride + kh
    fw = (wd - 1) * stride + kw
    full = np.zeros((n, ocg * groups, fh, fw))
    for g in range(groups):
        xs = x[:, g * icg:(g + 1) * icg]
        ws = w[g * icg:(g + 1) * icg]
        for ky in range(kh):
            for kx in range(kw):
                contrib = np.einsum('nihw,io->nohw', xs, ws[:, :, ky, kx])
                full[:, g * ocg:(g + 1) * ocg, ky:ky + (h - 1) * stride + 1:stride, kx:kx + (wd - 1) * stride + 1:stride] += contrib
    return full[:, :, py:fh - py, px:fw - px]
