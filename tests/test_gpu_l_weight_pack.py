"""GPU tests of the weight-operand packing kernel (pgpp_pack_weights, csrc/weight_pack.cu): plain, transposed, strided, split
precision, fp16, stacked (gamma | beta) and the polyphase up=2 form, checked by running the packed operand through the float64
emulation of the implicit GEMM (tests/helpers.py:emulate_igemm) against the oracle's convolutions; and the adjoint kernel of the
polyphase construction against autograd."""
import importlib

import pytest
import torch

from conftest import load_pkg
from helpers import emulate_igemm, rel_l2
from oracle import ref_ops

pytestmark = pytest.mark.gpu
load_pkg()
conv2d_gradfix = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
upfirdn2d = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
DEV = 'cuda:0'


@pytest.mark.parametrize('parts,tol', [(1, 8e-3), (2, 4e-5), (3, 3e-7)])
def test_weight_packing_plain_and_split_precision(parts, tol):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 20, 9, 11, generator=g)
    w = torch.randn(24, 20, 3, 3, generator=g)
    s = torch.randn(2, 20, generator=g) * 0.5 + 1
    for flip in (True, False):
        pw = conv2d_gradfix.packed_plain(w.to(DEV), flip, parts, 1, 1)
        assert pw.c_pad == 64 and pw.o_rows == 32 and pw.data.dtype == torch.bfloat16 and pw.data.shape[0] == parts
        want = ref_ops.conv2d_resample(x * s.reshape(2, 20, 1, 1), w, padding=1, flip_weight=flip)
        got = emulate_igemm(x, pw, scale=s)
        assert rel_l2(got, want) < tol     # bf16 expansion of the WEIGHTS only (activations stay exact here)


def test_weight_packing_transposed_layout_and_stride2():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 8, 10, 10, generator=g)
    wt = torch.randn(8, 6, 3, 3, generator=g)       # conv_transpose2d layout [I, O, kh, kw]
    pw = conv2d_gradfix.packed_plain(wt.to(DEV), False, 3, 2, 2, transpose_io=True)
    want = ref_ops.conv_transpose2d(x, wt, stride=1, padding=0)
    assert rel_l2(emulate_igemm(x, pw), want) < 3e-7
    w = torch.randn(6, 8, 3, 3, generator=g)
    pw2 = conv2d_gradfix.packed_plain(w.to(DEV), True, 3, 1, 1)
    assert rel_l2(emulate_igemm(x, pw2, stride=2), ref_ops.conv2d(x, w, stride=2, padding=1)) < 3e-7


def test_polyphase_up2_weights_reproduce_transposed_conv_plus_blur():
    g = torch.Generator().manual_seed(5)
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    x = torch.randn(2, 16, 8, 8, generator=g)
    w = torch.randn(32, 16, 3, 3, generator=g)
    for flipw in (False, True):
        pw = conv2d_gradfix.packed_up2(w.to(DEV), f.to(DEV), flipw, False, 3)
        assert pw.phases == 4 and pw.o == 32 and pw.o_rows == 128
        want = ref_ops.conv2d_resample(x, w, f=f, up=2, padding=1, flip_weight=flipw)
        got = emulate_igemm(x, pw)
        assert got.shape == want.shape == (2, 32, 16, 16)
        assert rel_l2(got, want) < 3e-7
    # asymmetric (non-separable-looking) filter exercises the flip conventions
    f2 = torch.rand(4, 4, generator=g)
    pw = conv2d_gradfix.packed_up2(w.to(DEV), f2.to(DEV), False, False, 3)
    assert rel_l2(emulate_igemm(x, pw), ref_ops.conv2d_resample(x, w, f=f2, up=2, padding=1, flip_weight=False)) < 3e-7
    pw = conv2d_gradfix.packed_up2(w.to(DEV), f2.to(DEV), False, True, 3)
    assert rel_l2(emulate_igemm(x, pw), ref_ops.conv2d_resample(x, w, f=f2, up=2, padding=1, flip_weight=False, flip_filter=True)) < 3e-7


def test_polyphase_weights_with_out_channels_not_multiple_of_16():
    g = torch.Generator().manual_seed(6)
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    x = torch.randn(1, 8, 6, 6, generator=g)
    w = torch.randn(24, 8, 3, 3, generator=g)
    pw = conv2d_gradfix.packed_up2(w.to(DEV), f.to(DEV), False, False, 3)
    assert pw.phase_stride == 32 and pw.o == 24 and pw.o_rows == 128
    assert rel_l2(emulate_igemm(x, pw), ref_ops.conv2d_resample(x, w, f=f, up=2, padding=1, flip_weight=False)) < 3e-7


def test_scale_master_rows_strided_source_and_stacked_tensors():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 12, 7, 9, generator=g)
    w_big = torch.randn(40, 24, 3, 3, generator=g).to(DEV)
    w = w_big[4:36:2, ::2]                                  # non-contiguous view: [16, 12, 3, 3]
    pw = conv2d_gradfix.packed_plain(w, True, 3, 1, 1, scale=0.37)
    assert rel_l2(emulate_igemm(x, pw), ref_ops.conv2d(x, w.cpu() * 0.37, padding=1)) < 3e-7
    # master rows = the float32 values the parts expand
    assert torch.equal(pw.master.reshape(9, pw.o_rows, pw.c_pad)[:, :16, :12].cpu(), (w.cpu() * 0.37).permute(2, 3, 0, 1).reshape(9, 16, 12))
    assert float(pw.master.reshape(9, pw.o_rows, pw.c_pad)[:, :, 12:].abs().max()) == 0.0
    # two tensors stacked along the output channels (gamma | beta of a SPADE block)
    wa, wb = torch.randn(16, 12, 3, 3, generator=g), torch.randn(16, 12, 3, 3, generator=g)
    pw2 = conv2d_gradfix.pack_weights_native([wa.to(DEV), wb.to(DEV)], 3, 3, 2, 1, 1, scale=0.5)
    assert pw2.o == 32 and pw2.o_rows == 32
    assert rel_l2(emulate_igemm(x, pw2), ref_ops.conv2d(x, torch.cat([wa, wb]) * 0.5, padding=1)) < 4e-5


def test_fp16_operand_and_other_source_dtypes():
    g = torch.Generator().manual_seed(8)
    w = torch.randn(16, 8, 3, 3, generator=g)
    for dt in (torch.float16, torch.bfloat16, torch.float64):
        pw = conv2d_gradfix.packed_plain(w.to(DEV, dt), True, 3, 1, 1)
        got = pw.data.float().sum(0).reshape(3, 3, pw.o_rows, pw.c_pad)[:, :, :16, :8].permute(2, 3, 0, 1).cpu()
        assert torch.allclose(got, w.to(dt).float(), rtol=1e-6, atol=1e-7)
    pw = conv2d_gradfix.packed_plain(w.to(DEV).half(), True, 1, 1, 1, f16=True)
    assert pw.data.dtype == torch.float16 and pw.f16 and pw.master is None
    got = pw.data[0].reshape(3, 3, pw.o_rows, pw.c_pad)[:, :, :16, :8].permute(2, 3, 0, 1).cpu()
    assert torch.equal(got, w.half())


def test_polyphase_adjoint_matches_autograd():
    """pgpp_up2_weight_adjoint = transpose of the linear map w -> Wp that pgpp_pack_weights applies"""
    custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
    plugin = custom_ops.get_plugin('conv2d_plugin')
    g = torch.Generator().manual_seed(9)
    f = torch.rand(4, 4, generator=g).to(DEV)
    o, ic = 6, 5
    for flipw in (False, True):
        for flipf in (False, True):
            gp = torch.randn(4, o, ic, 3, 3, generator=g).to(DEV)
            got = plugin.up2_weight_adjoint(gp, f, flipf, flipw, o, ic)
            # <gp, Wp(w)> differentiated w.r.t. w through the packing kernel's own forward map (probing with unit tensors)
            want = torch.zeros(o, ic, 3, 3)
            for ky in range(3):
                for kx in range(3):
                    e = torch.zeros(o, ic, 3, 3, device=DEV); e[:, :, ky, kx] = 1.0
                    pw = conv2d_gradfix.pack_weights_native(e, 3, 3, 3, 1, 1, flip=flipw, up2_filter=f, flip_filter=flipf)
                    wp = pw.master.reshape(3, 3, 4, pw.phase_stride, pw.c_pad)[:, :, :, :o, :ic].permute(2, 3, 4, 0, 1)   # [phase, o, i, a, b]
                    want[:, :, ky, kx] = (wp * gp).sum(dim=(0, 3, 4)).cpu()
            assert torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-6), (flipw, flipf)
