"""GPU parity tests for the weight-gradient kernel (pgpp_conv2d_wgrad) through conv2d_gradfix: what the reference gets
from cuDNN in Conv2dGradWeight.forward (torch_utils/ops/conv2d_gradfix.py:135-142) against float64 autograd of the
library convolution on the CPU (the oracle of a library call is the library's own definition)."""
import ctypes
import importlib

import pytest
import torch
import torch.nn.functional as F

from conftest import load_pkg
from helpers import rel_l2

pytestmark = pytest.mark.gpu
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'
TOL = {'bf16x3': 4e-5, 'bf16x2': 8e-5, 'bf16': 1.5e-2}


@pytest.fixture(autouse=True)
def _restore_precision():
    old = cg.fp32_precision
    yield
    cg.fp32_precision = old


def _reference(x, wshape, dy_seed, stride, pad, transpose, opad=0):
    wt = torch.zeros(*wshape, dtype=torch.float64, requires_grad=True)
    if transpose:
        y = F.conv_transpose2d(x.double(), wt, stride=stride, padding=pad, output_padding=opad)
    else:
        y = F.conv2d(x.double(), wt, stride=stride, padding=pad)
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(dy_seed))
    return dy, torch.autograd.grad(y, wt, dy.double())[0]


# n, o, i, h, w, k, stride, pad, transpose
CASES = [
    (2, 64, 64, 16, 64, 3, 1, 1, False),        # slab reuse, 16-wide K blocks
    (2, 32, 48, 8, 8, 3, 1, 1, False),          # slab reuse, 8-wide K blocks, ragged channels
    (3, 16, 16, 4, 4, 3, 1, 1, False),          # 4x4 images: several samples per K block
    (5, 16, 16, 1, 1, 3, 1, 1, False),          # 1x1 images
    (2, 128, 192, 32, 32, 1, 1, 0, False),      # 1x1 filter, two column blocks
    (1, 3, 64, 64, 64, 1, 1, 0, False),         # ToRGB
    (2, 64, 3, 40, 24, 7, 1, 3, False),         # 7x7 RGB stem
    (2, 64, 64, 32, 32, 3, 2, 1, False),        # stride 2
    (2, 64, 32, 33, 31, 3, 2, 0, False),        # stride 2 on the odd-size blurred image (conv2d_resample down=2)
    (1, 64, 64, 130, 260, 3, 1, 1, False),      # ragged tiles in both directions
    (2, 32, 64, 16, 16, 3, 2, 0, True),         # conv_transpose2d stride 2 (up=2 layer)
    (2, 64, 64, 16, 16, 3, 1, 1, True),         # conv_transpose2d stride 1
    (1, 260, 130, 16, 16, 3, 1, 1, False),      # several row blocks, ragged in both
]


@pytest.mark.parametrize('case', CASES, ids=[str(c) for c in CASES])
@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3', 'bf16'])
def test_weight_gradient_vs_float64_autograd(case, prec):
    n, o, i, h, w, k, s, p, tr = case
    x = torch.randn(n, i, h, w, generator=torch.Generator().manual_seed(41))
    wshape = (i, o, k, k) if tr else (o, i, k, k)
    dy, want = _reference(x, wshape, 42, s, p, tr)
    got = cg.weight_gradient(dy.to(DEV), x.to(DEV), wshape, s, (p, p), tr, precision=prec)
    assert got.dtype == torch.float32 and got.is_contiguous() and tuple(got.shape) == wshape
    assert rel_l2(got, want) < TOL[prec], rel_l2(got, want)


def test_weight_gradient_is_linear_and_matches_library_at_full_size():
    """Full-size layer (128 -> 128 channels, 3x3, 256 x 256, batch 8: the generator's dominant layer): compared with the
    library's fp32 weight gradient (TF32 off) and checked for linearity dW(a*dy1 + dy2) = a*dW(dy1) + dW(dy2)."""
    g = torch.Generator().manual_seed(43)
    x = torch.randn(8, 128, 256, 256, generator=g).to(DEV)
    dy1 = torch.randn(8, 128, 256, 256, generator=g).to(DEV)
    dy2 = torch.randn(8, 128, 256, 256, generator=g).to(DEV)
    wshape = (128, 128, 3, 3)
    g1 = cg.weight_gradient(dy1, x, wshape, 1, (1, 1), False, precision='bf16x2')
    g2 = cg.weight_gradient(dy2, x, wshape, 1, (1, 1), False, precision='bf16x2')
    g12 = cg.weight_gradient(0.5 * dy1 + dy2, x, wshape, 1, (1, 1), False, precision='bf16x2')
    assert rel_l2(g12, 0.5 * g1 + g2) < 1e-4
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        lib = torch.ops.aten.convolution_backward(dy1, x, torch.empty(wshape, device=DEV), None, [1, 1], [1, 1], [1, 1], False,
                                                  [0, 0], 1, [False, True, False])[1]
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel_l2(g1, lib) < 1e-4, rel_l2(g1, lib)


def test_weight_gradient_through_autograd_fp16_and_bf16_inputs():
    """half-precision activations (the reference's fp16 layers): dW comes back in the activation dtype"""
    g = torch.Generator().manual_seed(44)
    for dt in (torch.float16, torch.bfloat16):
        x = torch.randn(2, 32, 16, 16, generator=g).to(DEV, dt).requires_grad_(True)
        w = (torch.randn(48, 32, 3, 3, generator=g) * 0.1).to(DEV, dt).requires_grad_(True)
        y = cg.conv2d(x, w, padding=1)
        gw, = torch.autograd.grad(y.float().square().sum(), [w])
        xr, wr = x.detach().double().cpu().requires_grad_(True), w.detach().double().cpu().requires_grad_(True)
        yr = F.conv2d(xr, wr, padding=1)
        gr, = torch.autograd.grad(yr.square().sum(), [wr])
        assert gw.dtype == dt and rel_l2(gw.float(), gr) < 2e-2, rel_l2(gw.float(), gr)


@pytest.mark.parametrize('prec,tol', [('bf16x2', 1e-4), ('bf16x3', 5e-5)])
def test_conv_transpose_weight_gradient_through_autograd(prec, tol):
    cg.fp32_precision = prec
    g = torch.Generator().manual_seed(45)
    x0 = torch.randn(2, 24, 9, 11, generator=g)
    w0 = torch.randn(24, 20, 3, 3, generator=g) * 0.2
    for stride, pad, opad in [(1, 1, 0), (2, 1, 1), (2, 0, 0)]:
        xr, wr = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
        yr = F.conv_transpose2d(xr, wr, stride=stride, padding=pad, output_padding=opad)
        gr, = torch.autograd.grad(yr.square().sum(), [wr])
        xg, wg = x0.to(DEV).requires_grad_(True), w0.to(DEV).requires_grad_(True)
        yg = cg.conv_transpose2d(xg, wg, stride=stride, padding=pad, output_padding=opad)
        gg, = torch.autograd.grad(yg.square().sum(), [wg])
        assert rel_l2(gg, gr) < tol, (stride, pad, opad, rel_l2(gg, gr))


def test_wgrad_c_abi_rejects_bad_descriptors():
    lib = custom_ops.load_library()
    buf = torch.zeros(1 << 16, dtype=torch.bfloat16, device=DEV)
    out = torch.zeros(1 << 14, dtype=torch.float32, device=DEV)
    scratch = torch.zeros(1 << 14, dtype=torch.float32, device=DEV)

    def desc(**kw):
        d = custom_ops.WgradDesc()
        d.small = buf.data_ptr(); d.large = buf.data_ptr(); d.out = out.data_ptr(); d.workspace = scratch.data_ptr()
        d.s_parts = d.l_parts = 1
        d.n = 1; d.ca = 8; d.ca_pad = 64; d.hs = 8; d.ws = 8; d.cb = 8; d.cb_pad = 64; d.hl = 8; d.wl = 8
        d.kh = d.kw = 3; d.pad_y = d.pad_x = 1; d.stride = 1; d.products = 1
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    assert lib.pgpp_conv2d_wgrad(ctypes.byref(desc()), None) == 0
    for bad, msg in [(dict(stride=3), 'stride'), (dict(products=2), 'products'), (dict(ca_pad=48), 'ca_pad'),
                     (dict(products=3), 'parts'), (dict(workspace=0), 'workspace'), (dict(kh=9, kw=9), 'taps')]:
        assert lib.pgpp_conv2d_wgrad(ctypes.byref(desc(**bad)), None) != 0
        assert msg in lib.pgpp_last_error().decode(), (bad, lib.pgpp_last_error().decode())
    torch.cuda.synchronize()
