"""GPU parity tests for upfirdn2d (through the C ABI: pgpp_upfirdn2d)."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from helpers import max_abs, rel_l2, t
from oracle import ref_ops
from oracle.make_golden import UPFIRDN_CASES

pytestmark = pytest.mark.gpu
load_pkg()
upfirdn2d = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
DEV = 'cuda:0'


@pytest.mark.parametrize('case', UPFIRDN_CASES, ids=[c[0] for c in UPFIRDN_CASES])
def test_golden_forward_and_backward(case):
    name, shape, taps, sep, up, down, pad, flip, gain = case
    g = np.load(os.path.join(GOLDEN, 'upfirdn2d.npz'))
    x = t(g[f'{name}_x']).to(DEV).requires_grad_(True)
    f = t(g[f'{name}_f']).to(DEV) if g[f'{name}_f'].size else None
    y = upfirdn2d.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
    np.testing.assert_allclose(y.detach().cpu().numpy(), g[f'{name}_y'], rtol=1e-5, atol=3e-6)
    dx, = torch.autograd.grad(y, [x], t(g[f'{name}_dy']).to(DEV))
    np.testing.assert_allclose(dx.cpu().numpy(), g[f'{name}_dx'], rtol=1e-5, atol=3e-6)


def test_wrappers_golden():
    g = np.load(os.path.join(GOLDEN, 'upfirdn2d.npz'))
    x, f = t(g['wrap_x']).to(DEV), t(g['wrap_f']).to(DEV)
    np.testing.assert_allclose(upfirdn2d.filter2d(x, f).cpu().numpy(), g['wrap_filter2d'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(upfirdn2d.upsample2d(x, f).cpu().numpy(), g['wrap_upsample2d'], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(upfirdn2d.downsample2d(x, f).cpu().numpy(), g['wrap_downsample2d'], rtol=1e-5, atol=2e-6)


# the three families of the generator at reduced channel counts + ragged sizes
GEN_SHAPES = [
    ((2, 5, 65, 65), 1, 1, [1, 1, 1, 1], 4),        # blur after transposed conv (odd 2H+1 input)
    ((1, 3, 257, 257), 1, 1, [1, 1, 1, 1], 4),
    ((2, 4, 64, 64), 1, 1, [2, 2, 2, 2], 1),        # blur before strided conv
    ((1, 2, 300, 130), 1, 1, [2, 2, 2, 2], 1),      # ragged: not a multiple of the 128x32 tile
    ((2, 6, 64, 64), 1, 2, [1, 1, 1, 1], 1),        # 1x1-skip downsample
    ((2, 3, 32, 32), 2, 1, [2, 1, 2, 1], 4),        # image-skip upsample
    ((1, 1, 1, 1), 1, 1, [2, 2, 2, 2], 1),          # single pixel
]


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-6), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2), (torch.float64, 1e-12)])
@pytest.mark.parametrize('shape,up,down,pad,gain', GEN_SHAPES)
def test_generator_families_vs_oracle(shape, up, down, pad, gain, dtype, tol):
    g = torch.Generator().manual_seed(21)
    x = torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)
    f = ref_ops.setup_filter([1, 3, 3, 1])
    want = ref_ops.upfirdn2d(x.double(), f, up=up, down=down, padding=pad, gain=gain)
    got = upfirdn2d.upfirdn2d(x.to(DEV), f.to(DEV), up=up, down=down, padding=pad, gain=gain)
    assert got.dtype == dtype and tuple(got.shape) == tuple(want.shape)
    assert max_abs(got, want) <= tol * 8 * gain, max_abs(got, want)


# tiled up = 2 / down = 2 kernels: every parity of the padding (compile-time tap sets of fir_up2_kernel), negative padding (crop),
# non-square and asymmetric filters in both flip conventions, odd output widths (rows leave through the shared-memory stage)
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize('pad', [(2, 1, 2, 1), (1, 2, 2, 1), (2, 1, 1, 2), (1, 1, 1, 1), (3, 0, 0, 3), (0, 0, 0, 0), (-1, 2, 1, -2), (5, 4, 4, 5)], ids=str)
@pytest.mark.parametrize('shape', [(2, 3, 32, 32), (1, 2, 67, 131), (1, 5, 17, 300), (3, 1, 1, 1)], ids=str)
def test_up2_tiled_kernel_vs_oracle(shape, pad, dtype, tol):
    g = torch.Generator().manual_seed(23)
    x = torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)
    filters = [(ref_ops.setup_filter([1, 3, 3, 1]), False), (torch.randn(3, 4, generator=g), False), (torch.randn(4, 2, generator=g), True),
               (torch.randn(1, 1, generator=g), False)]
    for f, flip in filters:
        fh, fw = f.shape
        if shape[3] * 2 + pad[0] + pad[1] - fw + 1 < 1 or shape[2] * 2 + pad[2] + pad[3] - fh + 1 < 1:
            continue
        want = ref_ops.upfirdn2d(x.double(), f, up=2, padding=list(pad), flip_filter=flip, gain=4)
        got = upfirdn2d.upfirdn2d(x.to(DEV), f.to(DEV), up=2, padding=list(pad), flip_filter=flip, gain=4)
        assert got.dtype == dtype and tuple(got.shape) == tuple(want.shape)
        assert max_abs(got, want) <= tol * 8 * max(1.0, float(want.abs().max())), (fh, fw, flip, max_abs(got, want))


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize('shape,pad', [((2, 3, 66, 70), (1, 1, 1, 1)), ((1, 2, 131, 259), (2, 1, 1, 2)), ((1, 4, 37, 130), (0, 0, 0, 0)), ((1, 1, 9, 520), (1, 1, 1, 1))], ids=str)
def test_down2_tiled_kernel_ragged_and_odd_widths(shape, pad, dtype, tol):
    g = torch.Generator().manual_seed(24)
    x = torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)
    for f, flip in [(ref_ops.setup_filter([1, 3, 3, 1]), False), (torch.randn(4, 3, generator=g), True)]:
        want = ref_ops.upfirdn2d(x.double(), f, down=2, padding=list(pad), flip_filter=flip)
        got = upfirdn2d.upfirdn2d(x.to(DEV), f.to(DEV), down=2, padding=list(pad), flip_filter=flip)
        assert tuple(got.shape) == tuple(want.shape)
        assert max_abs(got, want) <= tol * 8 * max(1.0, float(want.abs().max())), max_abs(got, want)


def test_up2_is_the_adjoint_of_down2():
    """<down2(x), y> == <x, up2'(y)> with the padding list of the reference's backward (upfirdn2d.py:232-247): the property the D / R1
    backward relies on, at the full 512 px size of BASELINE configs[4]"""
    torch.manual_seed(3)
    f = ref_ops.setup_filter([1, 3, 3, 1]).to(DEV)
    x = torch.randn(2, 8, 512, 512, device=DEV, requires_grad=True)
    y = upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1])
    dy = torch.randn_like(y)
    dx, = torch.autograd.grad(y, [x], dy)
    lhs, rhs = float((y.double() * dy.double()).sum()), float((x.double() * dx.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs)) * 50, (lhs, rhs)
    want = upfirdn2d.upfirdn2d(dy, f, up=2, padding=[2, 1, 2, 1], flip_filter=True, impl='ref')
    assert float((dx - want).abs().max()) < 2e-5


# one-dimensional (separable) filters: the 12-tap sym6 passes of the ADA pipeline (augment.py:290,301) and other tap counts, up = 2 / down = 2 / neither
# along the filter axis, positive, zero and negative (cropping) padding, ragged sizes, both flip conventions
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.float16, 2e-3)])
@pytest.mark.parametrize('taps', [12, 8, 16, 5])
@pytest.mark.parametrize('shape', [(2, 3, 40, 52), (1, 2, 131, 259), (1, 1, 9, 7)], ids=str)
def test_separable_passes_vs_oracle(shape, taps, dtype, tol):
    g = torch.Generator().manual_seed(29 + taps)
    x = torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)
    f = torch.randn(taps, generator=g)
    for up, down, pad in [(2, 1, [taps // 2 + 1, taps // 2 - 1] * 2), (1, 2, [taps // 2 - 1, taps // 2] * 2), (1, 2, [-2, -1, -2, -1]), (1, 1, [3, 2, 0, 5]),
                          (2, 1, [0, 0, 0, 0])]:
        for flip in (False, True):
            if shape[3] * up + pad[0] + pad[1] - taps + 1 < down or shape[2] * up + pad[2] + pad[3] - taps + 1 < down:
                continue
            want = ref_ops.upfirdn2d(x.double(), f, up=up, down=down, padding=pad, flip_filter=flip, gain=up ** 2)
            got = upfirdn2d.upfirdn2d(x.to(DEV), f.to(DEV), up=up, down=down, padding=pad, flip_filter=flip, gain=up ** 2)
            assert got.dtype == dtype and tuple(got.shape) == tuple(want.shape), (up, down, pad)
            assert max_abs(got, want) <= tol * 8 * max(1.0, float(want.abs().max())), (up, down, pad, flip, max_abs(got, want))


def test_channels_last_and_strided_inputs():
    g = torch.Generator().manual_seed(22)
    x = torch.randn(2, 8, 20, 24, generator=g)
    f = ref_ops.setup_filter([1, 3, 3, 1])
    want = ref_ops.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], gain=4)
    xc = x.to(DEV).contiguous(memory_format=torch.channels_last)
    y = upfirdn2d.upfirdn2d(xc, f.to(DEV), up=2, padding=[2, 1, 2, 1], gain=4)
    assert y.is_contiguous(memory_format=torch.channels_last) and max_abs(y, want) < 3e-6
    big = torch.randn(2, 8, 20, 48, generator=g)
    view = big.to(DEV)[:, :, :, ::2]                                        # W stride 2
    assert max_abs(upfirdn2d.upfirdn2d(view, f.to(DEV), padding=[1, 1, 1, 1]), ref_ops.upfirdn2d(big[:, :, :, ::2], f, padding=[1, 1, 1, 1])) < 3e-6


def test_errors():
    x = torch.randn(1, 1, 4, 4, device=DEV)
    f = ref_ops.setup_filter([1, 3, 3, 1]).to(DEV)
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(x, f.double())                                   # f must be float32
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(x, f, padding=-3)                                # output smaller than 1x1
    with pytest.raises(AssertionError):
        upfirdn2d.upfirdn2d(x, f, up=0)


def test_full_size_blur_properties():
    # [4, 64, 513, 513] -> [4, 64, 512, 512]: the dominant upfirdn2d shape of the generator
    torch.manual_seed(0)
    f = ref_ops.setup_filter([1, 3, 3, 1]).to(DEV)
    x = torch.randn(4, 64, 513, 513, device=DEV)
    y = upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4)
    want = upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4, impl='ref')      # torch ops on the same device
    assert tuple(y.shape) == (4, 64, 512, 512) and float((y - want).abs().max()) < 2e-5
    # DC gain: a constant image stays constant (times gain) away from the border; linearity
    c = torch.full((1, 2, 513, 513), 0.5, device=DEV)
    yc = upfirdn2d.upfirdn2d(c, f, padding=[1, 1, 1, 1], gain=4)
    assert float((yc[:, :, 2:-2, 2:-2] - 2.0).abs().max()) < 1e-5
    a, b = torch.randn(1, 2, 513, 513, device=DEV), torch.randn(1, 2, 513, 513, device=DEV)
    lhs = upfirdn2d.upfirdn2d(a + 3 * b, f, padding=[1, 1, 1, 1])
    rhs = upfirdn2d.upfirdn2d(a, f, padding=[1, 1, 1, 1]) + 3 * upfirdn2d.upfirdn2d(b, f, padding=[1, 1, 1, 1])
    assert float((lhs - rhs).abs().max()) < 2e-5


@pytest.mark.parametrize('shape', [(2, 64, 32, 64), (1, 70, 17, 23), (3, 8, 5, 3), (2, 128, 64, 36)], ids=str)
@pytest.mark.parametrize('pad', [(2, 1, 2, 1), (1, 1, 1, 1), (2, 2, 2, 2)], ids=str)
def test_fir_pack_equals_blur_then_pack(shape, pad):
    """pgpp_fir_pack (blur written straight into the operand format, conv2d_resample.py:119-122) against the oracle's upfirdn2d
    on the CPU; with 3 bf16 parts the result carries 24 significand bits"""
    custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
    plugin = custom_ops.get_plugin('conv2d_plugin')
    n, c, h, w = shape
    g = torch.Generator().manual_seed(61)
    x = torch.randn(n, c, h, w, generator=g)
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    f_asym = torch.randn(3, 4, generator=g)                    # a non-symmetric, non-square filter checks the flip convention
    for filt, flip in ((f, False), (f_asym, False), (f_asym, True)):
        fh, fw = filt.shape
        want = ref_ops.upfirdn2d(x.double(), filt, padding=list(pad), flip_filter=flip, gain=1.7)
        c_pad = -(-c // 64) * 64
        for parts, tol in ((3, 2e-6), (2, 2e-5)):
            data = plugin.fir_pack(x.to(DEV), [float(v) for v in filt.reshape(-1)], fw, fh, *pad, flip, 1.7, c_pad, parts)
            assert tuple(data.shape) == (parts, n, want.shape[2], want.shape[3], c_pad)
            got = data.float().sum(0).permute(0, 3, 1, 2).cpu()
            assert torch.all(got[:, c:] == 0)
            assert max_abs(got[:, :c], want) <= tol * max(1.0, want.abs().max().item()), (fh, fw, flip, parts)
