"""GPU parity of the synthesis chain (the callers of the hot path) against the CPU oracle chain."""
import importlib

import pytest
import torch

from conftest import load_pkg
from helpers import max_abs, rel_l2
from oracle import ref_chain

pytestmark = pytest.mark.gpu
load_pkg()
synthesis = importlib.import_module('pgpp_b200.training.synthesis')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'


def _net(**kw):
    torch.manual_seed(0)
    net = synthesis.SynthesisChain(**kw).eval()
    for name, p in net.named_parameters():
        if name.endswith('noise_strength'):
            p.data.fill_(0.25)
        if name.endswith('bias') and 'affine' not in name:
            p.data.normal_(0, 0.5)
    return net


@pytest.mark.parametrize('prec,tol', [('bf16x3', 5e-5), ('bf16x2', 1e-4), ('bf16', 3e-2)])
def test_small_chain_fused_and_composition_routes(prec, tol):
    old = cg.fp32_precision
    cg.fp32_precision = prec
    try:
        net = _net(w_dim=64, img_resolution=64, channel_base=2048, channel_max=48, merge_channels=16)
        n = 3
        ws = torch.randn(n, net.num_ws, 64); pose = torch.randn(n, net.channels[8], 8, 8)
        cat = {'64': torch.randn(n, 16, 64, 64)}
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        want = ref_chain.synthesis_chain(sd, ws, pose, cat, img_resolution=64)
        net = net.to(DEV)
        catd = {k: v.to(DEV) for k, v in cat.items()}
        with torch.no_grad():
            for fused in (True, False):
                got = net(ws.to(DEV), pose.to(DEV), catd, fused=fused, noise_mode='const')
                for g, w, name in zip(got, want, ('img', 'parsing', 'texture')):
                    assert rel_l2(g, w) < tol, (prec, fused, name, rel_l2(g, w))
    finally:
        cg.fp32_precision = old


def test_full_size_512_chain_batch1_against_cpu_oracle():
    """the bench workload itself at batch 1: north-star tolerance, abs error relative to the output scale"""
    net = _net(w_dim=512, img_resolution=512)
    ws = torch.randn(1, net.num_ws, 512); pose = torch.randn(1, 512, 8, 8)
    cat = {str(r): torch.randn(1, 64, r, r).clamp_(-1, 1) for r in (64, 128, 256, 512)}
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    want = ref_chain.synthesis_chain(sd, ws, pose, cat, img_resolution=512)
    net = net.to(DEV)
    before = custom_ops.launch_count()
    with torch.no_grad():
        got = net(ws.to(DEV), pose.to(DEV), {k: v.to(DEV) for k, v in cat.items()}, noise_mode='const')
    assert custom_ops.launch_count() - before >= 24 * 3      # demod + pack + igemm per modulated conv
    for g, w, name in zip(got, want, ('img', 'parsing', 'texture')):
        assert tuple(g.shape) == tuple(w.shape)
        # end-to-end over ~15 chained layers (each <= 1e-4 per layer, tests/test_gpu_c_conv.py); measured ~7e-5..1e-4
        assert rel_l2(g, w) < 3e-4, (name, rel_l2(g, w))
        assert max_abs(g, w) <= 1e-3 * max(1.0, float(w.abs().max())), (name, max_abs(g, w), float(w.abs().max()))


@pytest.mark.parametrize('prec,tol', [('bf16x3', 5e-5), ('bf16x2', 1e-4), ('bf16', 3e-2)])
def test_operand_format_handover_route_small(prec, tol):
    """blocks >= PACKED_MIN_RES exchange bf16-split channels-innermost buffers (no packing pass, no concat, per-sample
    modulated weights); forced on at 32 px here so the small oracle-checkable net exercises it"""
    old, old_min = cg.fp32_precision, synthesis.PACKED_MIN_RES
    cg.fp32_precision = prec
    synthesis.PACKED_MIN_RES = 32
    try:
        net = _net(w_dim=64, img_resolution=128, channel_base=4096, channel_max=48, merge_channels=16)
        n = 3
        ws = torch.randn(n, net.num_ws, 64); pose = torch.randn(n, net.channels[8], 8, 8)
        cat = {'64': torch.randn(n, 16, 64, 64), '128': torch.randn(n, 16, 128, 128)}
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        want = ref_chain.synthesis_chain(sd, ws, pose, cat, img_resolution=128)
        net = net.to(DEV)
        with torch.no_grad():
            got = net(ws.to(DEV), pose.to(DEV), {k: v.to(DEV) for k, v in cat.items()}, fused=True, noise_mode='const')
        for g, w, name in zip(got, want, ('img', 'parsing', 'texture')):
            assert rel_l2(g, w) < tol, (prec, name, rel_l2(g, w))
    finally:
        cg.fp32_precision, synthesis.PACKED_MIN_RES = old, old_min


@pytest.mark.parametrize('prec,tol', [('bf16x2', 8e-5), ('bf16x3', 4e-5)])
def test_packed_activation_roundtrip_and_per_sample_weights(prec, tol):
    """one conv writing the operand format, a second one consuming it with style-modulated per-sample weights"""
    nets = importlib.import_module('pgpp_b200.training.networks')
    from oracle import ref_ops
    old = cg.fp32_precision
    cg.fp32_precision = prec
    try:
        g = torch.Generator().manual_seed(41)
        x = torch.randn(3, 32, 24, 20, generator=g); w1 = torch.randn(48, 32, 3, 3, generator=g) / 17
        w2 = torch.randn(32, 48, 3, 3, generator=g); s1 = torch.rand(3, 32, generator=g) + 0.5; s2 = torch.rand(3, 48, generator=g) + 0.5
        b1 = torch.randn(48, generator=g)
        y1 = ref_ops.synthesis_layer(x, s1, w1, b1, None, 1, None)
        y2 = ref_ops.modulated_conv2d(y1, w2, s2, padding=1)
        parts = cg._PRODUCTS[prec][1]
        buf = cg.PackedAct(torch.zeros(parts, 3, 24, 20, 128, dtype=torch.bfloat16, device=DEV), 48, 64)    # channels [64, 112) of a wider buffer
        with torch.no_grad():
            nets.modulated_conv2d_fused_act(x.to(DEV), w1.to(DEV), s1.to(DEV), padding=1, bias=b1.to(DEV), act='lrelu', clamp=256.0,
                                            out_packed=buf)
            assert rel_l2(buf.to_nchw(), y1) < tol
            got = nets.modulated_conv2d_fused_act(buf, w2.to(DEV), s2.to(DEV), padding=1)
        assert got.dtype == torch.float32 and rel_l2(got, y2) < tol
    finally:
        cg.fp32_precision = old


@pytest.mark.parametrize('shape', [(2, 128, 64, 48), (1, 70, 9, 13), (3, 64, 16, 128)], ids=str)
@pytest.mark.parametrize('terms', [1, 2])
def test_mix_pack_matches_the_masked_feature_composition(shape, terms):
    """pgpp_mix_pack = (x*(1-res) + mean*res) * mask, summed over branches (networks.py:2253-2276, 2307-2315), written in the
    operand format; with 0/1 masks every product is exact, so the packed sum equals the fp32 composition up to the bf16 split"""
    import importlib
    custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
    plugin = custom_ops.get_plugin('conv2d_plugin')
    n, c, h, w = shape
    g = torch.Generator().manual_seed(17)
    want = torch.zeros(n, c, h, w)
    args = []
    for t in range(terms):
        x = torch.randn(n, c, h, w, generator=g)
        m = torch.randn(n, c, 1, 1, generator=g)
        res = (torch.rand(n, 1, h, w, generator=g) > 0.7).float()
        mask = (torch.rand(n, 1, h, w, generator=g) > 0.4).float() if t == 0 else 1 - args[0][4]
        want += (x * (1 - res) + m * res) * mask
        args.append((x, m, (1 - res) * mask, res * mask, mask))
    c_pad = -(-c // 64) * 64
    for parts, tol in ((3, 1e-6), (2, 2e-5), (1, 4e-3)):
        data = plugin.mix_pack([tuple(v.to(DEV) for v in a[:4]) for a in args], c_pad, parts)
        assert tuple(data.shape) == (parts, n, h, w, c_pad)
        got = data.float().sum(0).permute(0, 3, 1, 2).cpu()
        assert torch.all(got[:, c:] == 0)
        assert (got[:, :c] - want).abs().max().item() <= tol * want.abs().max().item(), (parts, (got[:, :c] - want).abs().max().item())


@pytest.mark.parametrize('prec,tol', [('bf16x2', 8e-5), ('bf16x3', 5e-5), ('bf16', 2e-2)])
@pytest.mark.parametrize('n,ic,oc,h,w,noise_kind', [(2, 64, 64, 16, 16, 'const'), (3, 128, 32, 9, 20, 'per_sample'), (1, 64, 48, 32, 8, None)], ids=str)
def test_up2_layer_as_phase_gemms_plus_blur_pass(n, ic, oc, h, w, noise_kind, prec, tol):
    """The up = 2 synthesis layer on operand-format tensors: transposed convolution at 1x its MACs as four per-phase implicit GEMMs
    (conv2d_resample.py:125-139) + pgpp_fir_packed_act (blur, noise, bias, lrelu, gain, clamp) - against the CPU oracle's
    `synthesis_layer(up=2)` and against the one-launch polyphase form it replaces."""
    nets = importlib.import_module('pgpp_b200.training.networks')
    upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
    from oracle import ref_ops
    old, old_flag, old_min, old_t4 = cg.fp32_precision, nets.UP2_PHASES, nets.UP2_PHASES_MIN_IO, nets.UP2_TAPS4_MIN_IO
    cg.fp32_precision = prec
    nets.UP2_PHASES_MIN_IO = 0
    try:
        cg._init()
        g = torch.Generator().manual_seed(43)
        x = torch.randn(n, ic, h, w, generator=g); wt = torch.randn(oc, ic, 3, 3, generator=g)
        s = torch.rand(n, ic, generator=g) + 0.5; b = torch.randn(oc, generator=g)
        f = upf.setup_filter([1, 3, 3, 1])
        noise = None
        if noise_kind == 'const':
            noise = torch.randn(2 * h, 2 * w, generator=g) * 0.3
        elif noise_kind == 'per_sample':
            noise = torch.randn(n, 1, 2 * h, 2 * w, generator=g) * 0.3
        want = ref_ops.synthesis_layer(x, s, wt, b, noise, 2, f)
        parts = cg._PRODUCTS[prec][1]
        xp = cg.PackedAct(cg._plugin.pack_activations(x.to(DEV), None, -(-ic // 64) * 64, parts), ic)
        outs = {}
        for flag in (True, 'taps4', False):
            nets.UP2_PHASES = bool(flag)
            nets.UP2_PHASES_MIN_IO, nets.UP2_TAPS4_MIN_IO = ((1 << 30), 0) if flag == 'taps4' else (0, 1 << 30)
            out = cg.PackedAct(torch.zeros(parts, n, 2 * h, 2 * w, 128, dtype=torch.bfloat16, device=DEV), oc, 64)      # a channel slice of a wider buffer
            before = custom_ops.launch_count()
            with torch.no_grad():
                nets.modulated_conv2d_fused_act(xp, wt.to(DEV), s.to(DEV), noise=None if noise is None else noise.to(DEV), up=2, padding=1,
                                                resample_filter=f.to(DEV), flip_weight=False, bias=b.to(DEV), act='lrelu', clamp=256.0, out_packed=out)
            launches = custom_ops.launch_count() - before
            assert launches >= {True: 9, 'taps4': 3, False: 2}[flag], (flag, launches)          # demod + 4 x (modulate weights + GEMM) + blur pass  vs  demod + modulate + GEMM
            outs[flag] = out.to_nchw()
            assert torch.all(out.data[..., :64] == 0) and torch.all(out.data[..., 64 + oc:] == 0)       # neighbours of the slice untouched
            assert rel_l2(outs[flag], want) < tol, (prec, flag, rel_l2(outs[flag], want))
        assert rel_l2(outs[True], outs[False]) < 2 * tol and rel_l2(outs['taps4'], outs[False]) < 2 * tol
        nets.UP2_PHASES_MIN_IO, nets.UP2_TAPS4_MIN_IO = 0, 1 << 30
        # NCHW float32 input: packed once with the style scale folded in, phase GEMMs on shared weights
        nets.UP2_PHASES = True
        out = cg.PackedAct(torch.zeros(parts, n, 2 * h, 2 * w, 128, dtype=torch.bfloat16, device=DEV), oc, 64)
        with torch.no_grad():
            nets.modulated_conv2d_fused_act(x.to(DEV), wt.to(DEV), s.to(DEV), noise=None if noise is None else noise.to(DEV), up=2, padding=1,
                                            resample_filter=f.to(DEV), flip_weight=False, bias=b.to(DEV), act='lrelu', clamp=256.0, out_packed=out)
        assert rel_l2(out.to_nchw(), want) < tol
        # float32 NCHW output (the low-resolution blocks hand tensors over), tensor and operand-format input
        for src in (x.to(DEV), xp):
            with torch.no_grad():
                y = nets.modulated_conv2d_fused_act(src, wt.to(DEV), s.to(DEV), noise=None if noise is None else noise.to(DEV), up=2, padding=1,
                                                    resample_filter=f.to(DEV), flip_weight=False, bias=b.to(DEV), act='lrelu', clamp=256.0)
            assert y.dtype == torch.float32 and y.is_contiguous() and tuple(y.shape) == tuple(want.shape) and rel_l2(y, want) < tol
    finally:
        cg.fp32_precision, nets.UP2_PHASES, nets.UP2_PHASES_MIN_IO, nets.UP2_TAPS4_MIN_IO = old, old_flag, old_min, old_t4
