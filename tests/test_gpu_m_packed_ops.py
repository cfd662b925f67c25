"""GPU parity tests of the operand-format (packed -> packed) companions of the convolution kernel: pgpp_fir_packed (the FIR in
front of a down-sampling convolution, conv2d_resample.py:107-122, run on the bf16 expansion), the accumulate mode of the lean
operand-format epilogue (residual adds without leaving the format), and the encoder chains built from them."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_pkg
from helpers import rel_l2
from oracle import ref_ops

pytestmark = pytest.mark.gpu
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
syn = importlib.import_module('pgpp_b200.training.synthesis')
gen = importlib.import_module('pgpp_b200.training.generator')
DEV = 'cuda:0'
# the operand format itself carries 8 * parts significand bits: error of one round trip through it
FMT_TOL = {1: 4e-3, 2: 2e-5, 3: 2e-7}


@pytest.fixture(autouse=True)
def _restore_precision():
    old = cg.fp32_precision
    yield
    cg.fp32_precision = old


FIR_CASES = [
    # name,             n, c,   h,  w,  filter,        down, padding (x0, x1, y0, y1), gain
    ('blur_even',       2, 64,  32, 48, [1, 3, 3, 1],  1,    (2, 2, 2, 2), 1.0),
    ('blur_odd_ragged', 3, 24,  17, 29, [1, 3, 3, 1],  1,    (2, 2, 2, 2), 1.0),
    ('blur_after_up',   1, 128, 33, 33, [1, 3, 3, 1],  1,    (1, 1, 1, 1), 4.0),
    ('down2_skip',      2, 64,  32, 32, [1, 3, 3, 1],  2,    (1, 1, 1, 1), 1.0),
    ('down2_odd',       2, 72,  19, 23, [1, 3, 3, 1],  2,    (1, 1, 1, 1), 1.0),
    ('taps3',           2, 16,  12, 20, [1, 2, 1],     1,    (1, 1, 1, 1), 1.0),
    ('taps2_asym_pad',  2, 8,   9,  9,  [1, 1],        2,    (1, 0, 0, 1), 2.0),
    ('one_pixel_out',   1, 8,   2,  2,  [1, 3, 3, 1],  1,    (1, 1, 1, 1), 1.0),
    ('tall_strips',     1, 8,   67, 5,  [1, 3, 3, 1],  1,    (2, 2, 2, 2), 1.0),
]


@pytest.mark.parametrize('parts', [1, 2, 3])
@pytest.mark.parametrize('flip', [False, True])
@pytest.mark.parametrize('case', FIR_CASES, ids=[c[0] for c in FIR_CASES])
def test_fir_packed_vs_oracle(case, flip, parts):
    name, n, c, h, w, taps, down, pad, gain = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    x = torch.randn(n, c, h, w, generator=g)
    f = ref_ops.setup_filter(taps)
    if len(taps) == 3:
        f = f * torch.tensor([[1.0, 2.0, 0.5]])      # asymmetric filter: flip_filter must matter
    prec = {1: 'bf16', 2: 'bf16x2', 3: 'bf16x3'}[parts]
    xp = cg.pack_operand(x.to(DEV), prec)
    x_seen = xp.to_nchw().cpu()                      # what the kernel reads: the bf16 expansion of x
    want = ref_ops.upfirdn2d(x_seen, f, down=down, padding=list(pad), flip_filter=flip, gain=gain)
    got = cg.fir_packed(xp, f.to(DEV), down=down, padding=pad, flip_filter=flip, gain=gain)
    assert tuple(got.shape) == tuple(want.shape)
    assert rel_l2(got.to_nchw(), want) < FMT_TOL[parts], (name, rel_l2(got.to_nchw(), want))
    # padding channels of a freshly allocated result are zero (they meet zero weights but must not be NaN / Inf patterns)
    assert torch.isfinite(got.data.float()).all()


def test_fir_packed_slice_copy_and_into():
    """f=None is the channel-slice copy the synthesis blocks use to place the garment features next to conv1's output"""
    g = torch.Generator().manual_seed(5)
    a = torch.randn(2, 64, 16, 24, generator=g)
    src = cg.pack_operand(a.to(DEV), 'bf16x2')
    buf = cg.PackedAct.empty(2, 16, 24, 128 + 64, 2, DEV)
    buf.zero_()
    cg.fir_packed(src, None, out=cg.PackedAct(buf, 64, 128))
    whole = cg.PackedAct(buf, 192, 0).to_nchw()
    assert torch.equal(whole[:, 128:], src.to_nchw())
    assert float(whole[:, :128].abs().max()) == 0.0
    # blur out of a slice of a wider buffer into a slice of another one
    f = ref_ops.setup_filter([1, 3, 3, 1])
    dst = cg.PackedAct.empty(2, 17, 25, 128, 2, DEV)
    dst.zero_()
    cg.fir_packed(cg.PackedAct(buf, 64, 128), f.to(DEV), padding=(2, 2, 2, 2), out=cg.PackedAct(dst, 64, 64))
    want = ref_ops.upfirdn2d(src.to_nchw().cpu(), f, padding=[2, 2, 2, 2])
    assert rel_l2(cg.PackedAct(dst, 64, 64).to_nchw(), want) < FMT_TOL[2]
    assert float(cg.PackedAct(dst, 64, 0).to_nchw().abs().max()) == 0.0


@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3', 'bf16'])
@pytest.mark.parametrize('shape', [(2, 64, 64, 32, 32, 3), (2, 128, 128, 16, 40, 3), (3, 64, 128, 24, 24, 1), (2, 64, 64, 128, 64, 3)],
                         ids=['64x64', '128x128', '1x1_64_128', 'resident_slab'])
def test_packed_accumulate(shape, prec):
    """out_packed + accumulate: y(packed) += act(conv(x) + b) * gain, the residual add of ResBlock / Spade_ResBlockV4_512 on the format"""
    n, ic, oc, h, w, k = shape
    cg.fp32_precision = prec
    parts = cg._PRODUCTS[prec][1]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(n, ic, h, w, generator=g)
    wt = torch.randn(oc, ic, k, k, generator=g) / np.sqrt(ic * k * k)
    b = torch.randn(oc, generator=g)
    y0 = torch.randn(n, oc, h, w, generator=g)
    xp = cg.pack_operand(x.to(DEV), prec)
    yp = cg.pack_operand(y0.to(DEV), prec)
    y_seen = yp.to_nchw().cpu().double()
    pw = cg.packed_plain(wt.to(DEV), True, parts, k // 2, k // 2)
    cg.igemm_conv(xp, pw, bias=b.to(DEV), act='relu', gain=0.75, out_packed=yp, accumulate=True)
    conv = torch.nn.functional.conv2d(xp.to_nchw().cpu().double(), wt.double(), b.double(), padding=k // 2)
    want = y_seen + conv.clamp_min(0) * 0.75
    tol = {'bf16x2': 8e-5, 'bf16x3': 4e-5, 'bf16': 1.5e-2}[prec]
    assert rel_l2(yp.to_nchw(), want) < tol, rel_l2(yp.to_nchw(), want)


@pytest.mark.parametrize('k', [1, 3])
def test_down_conv_on_packed_input_matches_tensor_route(k):
    """Conv2dLayer(down=2) fed a PackedAct (FIR on the operand format) against the same layer fed the tensor (float32 FIR, then packing)"""
    torch.manual_seed(3)
    layer = syn.Conv2dLayer(64, 128, kernel_size=k, down=2, activation='lrelu').to(DEV).eval()
    layer.bias.data.normal_()
    x = torch.randn(2, 64, 64, 48, device=DEV)
    with torch.no_grad():
        want = layer(x, fused=True)
        xp = cg.pack_operand(x, cg.fp32_precision)
        out = cg.PackedAct(cg.PackedAct.empty(2, 32, 24, 128, syn._parts(), DEV), 128)
        layer(xp, fused=True, out_packed=out)
        comp = layer(x, fused=False)         # conv2d_resample composition (reference call sequence)
    assert rel_l2(out.to_nchw(), want) < 3e-5
    assert rel_l2(out.to_nchw(), comp) < 8e-5


def test_resblock_packed_output_matches_tensor_output():
    torch.manual_seed(4)
    blk = gen.ResBlock(64, 64, kernel_size=4, activation='relu').to(DEV).eval()
    blk2 = gen.ResBlock(64, 128, kernel_size=4, activation='relu', down=2).to(DEV).eval()
    x = torch.randn(2, 64, 48, 64, device=DEV)
    with torch.no_grad():
        want = blk(x, fused=True)
        got = blk(x, fused=True, out_packed=True)
        assert isinstance(got, cg.PackedAct)
        assert rel_l2(got.to_nchw(), want) < 3e-5
        want2 = blk2(want, fused=True)
        got2 = blk2(got, fused=True)                # operand-format input into the down-sampling block
        comp2 = blk2(blk(x, fused=False), fused=False)
    assert rel_l2(got2, want2) < 5e-5
    assert rel_l2(got2, comp2) < 1e-4


def test_const_encoder_packed_chain_matches_composition():
    torch.manual_seed(6)
    enc = gen.ConstEncoderNetwork(input_nc=5, output_nc=512, ngf=64, n_downsampling=6).to(DEV).eval()
    x = torch.randn(2, 5, 128, 128, device=DEV).clamp(-1, 1)
    with torch.no_grad():
        got = enc(x, fused=True)
        want = enc(x, fused=False)
    assert torch.is_tensor(got) and tuple(got.shape) == tuple(want.shape) == (2, 512, 2, 2)
    assert rel_l2(got, want) < 1e-4, rel_l2(got, want)


@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3', 'bf16'])
@pytest.mark.parametrize('cfg', [(64, 7, True, 40, 56), (128, 0, True, 32, 32), (512, 0, False, 16, 24), (64, 7, False, 17, 19)],
                         ids=['c64_parsing_acc', 'c128_acc', 'c512_noacc', 'ragged_noacc'])
def test_torgb_thin_kernel(cfg, prec):
    """ToRGB (+ parsing head) on an operand-format input: the thin CUDA-core kernel against the tensor-core route and float64"""
    ic, pc, acc, h, w = cfg
    cg.fp32_precision = prec
    torch.manual_seed(9)
    layer = syn.ToRGBLayer(ic, 3, w_dim=32, conv_clamp=2.0, parsing_channels=pc).to(DEV).eval()
    layer.bias.data.normal_()
    if pc:
        layer.m_bias1.data.normal_()
    x = torch.randn(3, ic, h, w, device=DEV)
    ws = torch.randn(3, 32, device=DEV)
    img0 = torch.randn(3, 3, h, w, device=DEV)
    xp = cg.pack_operand(x, prec)
    x_seen = xp.to_nchw().double()
    with torch.no_grad():
        got, got_pp = layer(xp, ws, img=img0.clone() if acc else None)
        syn.THIN_TORGB = False
        try:
            ref, ref_pp = layer(xp, ws, img=img0.clone() if acc else None)
        finally:
            syn.THIN_TORGB = True
        styles = (layer.affine(ws) * layer.weight_gain).double()
        want = torch.einsum('nchw,oc,nc->nohw', x_seen, layer.weight.double().reshape(3, ic), styles) + layer.bias.double().reshape(1, 3, 1, 1)
        want = want.clamp(-2, 2) + (img0.double() if acc else 0)
    # the thin kernel multiplies the exact float32 weights: its only error is float32 accumulation
    assert rel_l2(got, want) < 2e-6, rel_l2(got, want)
    assert rel_l2(got, ref) < {'bf16x2': 8e-5, 'bf16x3': 4e-5, 'bf16': 1.5e-2}[prec]
    if pc:
        want_pp = (torch.einsum('nchw,oc,nc->nohw', x_seen, layer.m_weight1.double().reshape(pc, ic), styles) + layer.m_bias1.double().reshape(1, pc, 1, 1)).clamp(-2, 2)
        assert rel_l2(got_pp, want_pp) < 2e-6
        assert rel_l2(got_pp, ref_pp) < {'bf16x2': 8e-5, 'bf16x3': 4e-5, 'bf16': 1.5e-2}[prec]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize('shape', [(2, 64, 16, 32), (1, 70, 9, 20), (2, 24, 7, 13), (1, 128, 4, 130), (3, 8, 33, 6)], ids=str)
def test_pack_activations_every_source_dtype_and_alignment(shape, dtype):
    """pgpp_pack_activations (the entry into the operand format): float32 / float16 / bfloat16 NCHW sources, widths that are and are not
    multiples of 4 (128-bit / 64-bit vector loads vs the scalar path), a row-strided view, with and without the style scale; the sum
    of the 2 bf16 parts reproduces the source to 16 significand bits, padding channels are zero."""
    cg._init()
    n, c, h, w = shape
    g = torch.Generator().manual_seed(71)
    x = torch.randn(n, c, h, w, generator=g).to(dtype).to(DEV)
    s = torch.randn(n, c, generator=g).to(DEV)
    c_pad = -(-c // 64) * 64
    views = [x]
    if w % 4 == 0:
        big = torch.randn(n, c, h, w + 4, generator=g).to(dtype).to(DEV)
        views.append(big[..., :w])                                      # row pitch w + 4: still 4-element aligned rows
    views.append(torch.randn(n, c, h, w + 1, generator=g).to(dtype).to(DEV)[..., 1:])      # misaligned base: scalar path
    for src in views:
        for scale in (None, s):
            data = cg._plugin.pack_activations(src, scale, c_pad, 2)
            assert tuple(data.shape) == (2, n, h, w, c_pad)
            got = data.float().sum(0).permute(0, 3, 1, 2)
            want = src.float() * (1 if scale is None else scale[:, :, None, None])
            assert torch.all(got[:, c:] == 0)
            err = float((got[:, :c] - want).abs().max())
            assert err <= 2e-5 * max(1.0, float(want.abs().max())), (err, tuple(src.stride()))
    if dtype == torch.float16:                                          # native f16 operand (one part, exact copy)
        data = cg._plugin.pack_activations(x, None, c_pad, 1, f16=True)
        got = data.view(torch.float16)[0].permute(0, 3, 1, 2)[:, :c]
        assert torch.equal(got, x)


@pytest.mark.parametrize('prec', ['bf16x2', 'bf16'])
@pytest.mark.parametrize('n,ic,oc,h,w', [(2, 64, 64, 32, 32), (3, 128, 128, 16, 32), (1, 64, 128, 64, 16), (2, 64, 48, 16, 16), (2, 64, 64, 20, 24)], ids=str)
def test_instance_norm_statistics_from_the_conv_epilogue(n, ic, oc, h, w, prec):
    """igemm_conv(instnorm_eps=...) returns the statistics `torch.nn.InstanceNorm2d` / `torch.var_mean(y, (2, 3), unbiased=False)` would
    compute on the convolution's OWN float32 output (networks.py:1702-1723 normalises what the preceding conv wrote): warp partials
    around a pivot in the epilogue + a float64 merge.  Checked against float64 statistics of the returned tensor, including a channel
    with a large mean and a tiny variance (cancellation) and a constant channel; (20, 24) is not a multiple of the pixel tile and takes
    the torch.var_mean fall-back."""
    cg._init()
    cg.fp32_precision = prec
    g = torch.Generator().manual_seed(81)
    x = torch.randn(n, ic, h, w, generator=g).to(DEV)
    wt = torch.randn(oc, ic, 3, 3, generator=g).to(DEV) / (ic * 9) ** 0.5
    bias = torch.randn(oc, generator=g).to(DEV)
    wt[1] = 0; bias[1] = 3.0                       # constant channel: variance exactly 0
    wt[2] *= 1e-4; bias[2] = 50.0                  # mean^2 / var ~ 1e10
    parts = cg._PRODUCTS[prec][1]
    xp = cg.PackedAct(cg._plugin.pack_activations(x, None, ic, parts), ic)
    pw = cg.packed_plain(wt, True, parts, 1, 1)
    y, mean, rstd = cg.igemm_conv(xp, pw, bias=bias, instnorm_eps=1e-5)
    assert tuple(y.shape) == (n, oc, h, w) and tuple(mean.shape) == (n, oc) and tuple(rstd.shape) == (n, oc)
    y_plain = cg.igemm_conv(xp, pw, bias=bias)
    assert torch.equal(y, y_plain)                                      # the statistics do not change what is written
    yd = y.double()
    want_mean = yd.mean(dim=(2, 3))
    want_var = yd.var(dim=(2, 3), unbiased=False)
    want_rstd = (want_var + 1e-5).rsqrt()
    assert float((mean.double() - want_mean).abs().max()) <= 2e-6 * max(1.0, float(want_mean.abs().max()))
    assert float(((rstd.double() - want_rstd) / want_rstd).abs().max()) <= 2e-5
    assert float(mean[:, 1].sub(3.0).abs().max()) == 0.0 and float((rstd[:, 1] - 1e-5 ** -0.5).abs().max()) <= 1e-2
    # and against the library statistics the unfused route uses
    var_t, mean_t = torch.var_mean(y, dim=(2, 3), unbiased=False)
    assert float((mean - mean_t).abs().max()) <= 1e-5 * max(1.0, float(mean_t.abs().max()))
    assert float(((rstd - (var_t + 1e-5).rsqrt()) / want_rstd.float()).abs().max()) <= 1e-3


def test_spade_block_with_fused_statistics_equals_the_var_mean_route():
    cg._init()
    torch.manual_seed(5)
    blk = gen.Spade_ResBlockV4_512(64, 64, spade_channels=1).to(DEV).eval()
    x = torch.randn(2, 64, 64, 64, device=DEV)
    parsing = torch.randint(0, 7, (2, 1, 64, 64), device=DEV).float()
    xp = lambda: cg.PackedAct(cg._plugin.pack_activations(x, None, 64, 2), 64)
    with torch.no_grad():
        old = gen.FUSE_INSTNORM_STATS
        try:
            gen.FUSE_INSTNORM_STATS = True
            a = blk(xp(), parsing)
            gen.FUSE_INSTNORM_STATS = False
            b = blk(xp(), parsing)
        finally:
            gen.FUSE_INSTNORM_STATS = old
    assert rel_l2(a, b) < 2e-6


@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3'])
def test_spade_norm_block_training_route_with_fused_relu(prec):
    """training route of Spade_Norm_Block (networks.py:1702-1723) with conv_mlp's ReLU in the convolution's epilogue - output and every gradient
    against the route with nn.ReLU as its own op (generator.FUSE_MLP_RELU = False) and against float64 of the same module on the CPU (library
    convolutions)"""
    import copy
    cg.fp32_precision = prec
    torch.manual_seed(77)
    blk = gen.Spade_Norm_Block(24, 32).to(DEV)
    x0 = torch.randn(2, 32, 20, 28, device=DEV)
    f0 = torch.randn(2, 24, 20, 28, device=DEV)
    probe = torch.randn(2, 32, 20, 28, device=DEV)

    def run(module, x0_, f0_, probe_, **kw):
        x, f = x0_.clone().requires_grad_(True), f0_.clone().requires_grad_(True)
        y = module(x, f, **kw)
        grads = torch.autograd.grad((y * probe_).sum() + 0.1 * y.square().sum(), [x, f] + list(module.parameters()))
        return [y] + list(grads)

    try:
        fused_relu = run(blk, x0, f0, probe, fused=False)
        gen.FUSE_MLP_RELU = False
        plain = run(blk, x0, f0, probe, fused=False)
    finally:
        gen.FUSE_MLP_RELU = True
    ref = copy.deepcopy(blk).double().cpu()
    for m in ref.modules():         # the FIR taps stay float32, as upfirdn2d.setup_filter leaves them (upfirdn2d.py:91)
        if isinstance(getattr(m, 'resample_filter', None), torch.Tensor):
            m.resample_filter = m.resample_filter.float()
    want = run(ref, x0.double().cpu(), f0.double().cpu(), probe.double().cpu(), fused=False, impl='ref')
    tol = {'bf16x2': 2e-4, 'bf16x3': 1e-4}[prec]
    assert len(fused_relu) == len(plain) == len(want)
    for i, (a, b, r) in enumerate(zip(fused_relu, plain, want)):
        assert a.shape == b.shape and rel_l2(a, b) < tol, (i, rel_l2(a, b))
        # gradients that pass the ReLU gate (denorm_feats, conv_mlp.weight) see it flip where the pre-activation is within rounding of zero:
        # a 1e-5 fraction of the pixels, an O(1) change each - in any finite precision
        assert rel_l2(a, r) < (tol if i < 2 else 1e-2), (i, rel_l2(a, r))
