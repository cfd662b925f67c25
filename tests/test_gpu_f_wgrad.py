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


# n, o, i, h, w, (kh, kw), (pad_y, pad_x), transpose, f16
PAIR_CASES = [
    (2, 64, 64, 40, 36, (3, 3), (1, 1), False, False),      # the 64 -> 64 layers: two taps per accumulator, three kx per tile
    (2, 24, 20, 13, 11, (3, 3), (1, 1), False, False),      # ragged channels and tiles, 8-wide K blocks
    (1, 64, 192, 20, 48, (3, 3), (1, 1), False, False),     # three column blocks of the larger operand
    (2, 48, 64, 17, 33, (3, 3), (0, 0), False, False),      # valid convolution: L has more rows / columns than S
    (2, 64, 64, 16, 16, (3, 3), (2, 2), False, False),      # 'full' padding
    (2, 32, 32, 24, 16, (3, 1), (1, 0), False, False),      # one filter column
    (2, 32, 32, 24, 16, (2, 2), (1, 0), False, False),      # even filter: one tap pair, no idle half
    (2, 32, 32, 24, 20, (4, 2), (1, 1), False, False),      # two full pairs
    (2, 16, 64, 24, 20, (3, 4), (1, 2), False, False),      # four kx: 256 columns
    (2, 64, 64, 24, 24, (3, 3), (1, 1), True, False),       # conv_transpose2d: S is the input
    (2, 64, 64, 40, 36, (3, 3), (1, 1), False, True),       # f16 operands (fp16 layers of the discriminator)
    (8, 8, 8, 8, 8, (3, 3), (1, 1), False, False),          # smallest image that takes the pair form
]


@pytest.mark.parametrize('case', PAIR_CASES, ids=[str(c) for c in PAIR_CASES])
@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3'])
def test_weight_gradient_tap_pair_form(case, prec):
    """at most 64 channels on the accumulator-lane side: two vertical taps share the 128 lanes (the A descriptor's leading byte offset is one
    row of the S slab) and the horizontal taps sit side by side in the tile - against float64 autograd and against the one-tap-per-accumulator
    form of the same kernel (PGPP_WGRAD_NO_PAIR)"""
    import os
    n, o, i, h, w, (kh, kw), (py, px), tr, f16 = case
    if f16 and prec != 'bf16x2':
        pytest.skip('f16 operands have one precision')
    g = torch.Generator().manual_seed(45)
    x = torch.randn(n, i, h, w, generator=g)
    wshape = (i, o, kh, kw) if tr else (o, i, kh, kw)
    wt = torch.zeros(*wshape, dtype=torch.float64, requires_grad=True)
    dt = torch.float16 if f16 else torch.float32
    x = x.to(dt)
    y = F.conv_transpose2d(x.double(), wt, padding=(py, px)) if tr else F.conv2d(x.double(), wt, padding=(py, px))
    dy = torch.randn(y.shape, generator=g).to(dt)
    want = torch.autograd.grad(y, wt, dy.double())[0]
    kw_ = dict(precision='f16' if f16 else prec, out_dtype=torch.float32)
    got = cg.weight_gradient(dy.to(DEV), x.to(DEV), wshape, 1, (py, px), tr, **kw_)
    os.environ['PGPP_WGRAD_NO_PAIR'] = '1'
    custom_ops.refresh_env()
    try:
        old = cg.weight_gradient(dy.to(DEV), x.to(DEV), wshape, 1, (py, px), tr, **kw_)
    finally:
        del os.environ['PGPP_WGRAD_NO_PAIR']
        custom_ops.refresh_env()
    tol = 1e-5 if f16 else TOL[prec]      # f16 operands are exact inputs here: only the fp32 accumulation order differs
    assert tuple(got.shape) == wshape and rel_l2(got.float(), want) < tol, rel_l2(got.float(), want)
    assert rel_l2(got.float(), old.float()) < tol, rel_l2(got.float(), old.float())


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
