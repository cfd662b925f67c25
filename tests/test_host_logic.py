"""CPU tests (-m "not gpu"): the drop-in modules' PyTorch ('ref') paths and host-side logic against the
golden fixtures minted from the reference, the operand packing against the oracle, and the C-ABI surface."""
import ctypes
import importlib
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, load_pkg
from helpers import emulate_igemm, rel_l2, t, upfirdn2d_ref_on_cpu
from oracle import ref_ops
from oracle.make_golden import CONV_CASES, MODCONV_CASES, UPFIRDN_CASES

pkg = load_pkg()
ops = lambda name: importlib.import_module(f'pgpp_b200.torch_utils.ops.{name}')
bias_act = ops('bias_act'); upfirdn2d = ops('upfirdn2d'); conv2d_resample = ops('conv2d_resample')
conv2d_gradfix = ops('conv2d_gradfix'); fma = ops('fma')
networks = importlib.import_module('pgpp_b200.training.networks')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')


def _opt(v):
    return None if v == 'None' else float(v)


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'pgpp.h')).read()
    declared = set(re.findall(r'PGPP_API\s+[\w\s\*]+?\b(pgpp_\w+)\s*\(', header))
    assert declared == set(custom_ops.EXPORTED_SYMBOLS), declared ^ set(custom_ops.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(custom_ops.library_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert custom_ops.load_library().pgpp_version() >= 100
    # struct mirror has the size the C compiler gives pgpp_conv_desc (checked against a tiny C program's sizeof)
    assert ctypes.sizeof(custom_ops.ConvDesc) % 8 == 0


def test_cuda_impl_never_falls_back_on_cpu_tensors():
    x = torch.randn(2, 3, 4, 4)
    with pytest.raises(RuntimeError):
        bias_act.bias_act(x, torch.randn(3))
    with pytest.raises(RuntimeError):
        upfirdn2d.upfirdn2d(x, upfirdn2d.setup_filter([1, 3, 3, 1]))
    with pytest.raises(AssertionError):
        bias_act.bias_act(x, impl='fast')
    assert not conv2d_gradfix._should_use_custom_op(x)


def test_get_plugin_contract():
    p = custom_ops.get_plugin('bias_act_plugin', sources=['ignored.cu'], extra_cuda_cflags=['--use_fast_math'])
    assert p is custom_ops.get_plugin('bias_act_plugin')
    assert hasattr(p, 'bias_act') and hasattr(custom_ops.get_plugin('upfirdn2d_plugin'), 'upfirdn2d')
    with pytest.raises(RuntimeError):
        custom_ops.get_plugin('nope_plugin')
    assert custom_ops.verbosity in ('none', 'brief', 'full')


def test_bias_act_ref_path_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, 'bias_act.npz'))
    x, b = t(g['x']), t(g['b'])
    assert list(bias_act.activation_funcs) == list(ref_ops.ACTIVATIONS)
    for name, spec in bias_act.activation_funcs.items():
        da, dg, idx, ref, has2 = ref_ops.ACTIVATIONS[name]
        assert (spec.def_alpha, float(spec.def_gain), spec.cuda_idx, spec.ref, spec.has_2nd_grad) == (da, float(dg), idx, ref, has2)
    n = len([k for k in g.files if k.endswith('_meta')])
    for i in range(n):
        act, alpha, gain, clamp = g[f'case{i}_meta']
        y = bias_act.bias_act(x, b, act=str(act), alpha=_opt(alpha), gain=_opt(gain), clamp=_opt(clamp), impl='ref')
        np.testing.assert_allclose(y.numpy(), g[f'case{i}_y'], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('case', UPFIRDN_CASES, ids=[c[0] for c in UPFIRDN_CASES])
def test_upfirdn2d_ref_path_matches_reference_golden(case):
    name, shape, taps, sep, up, down, pad, flip, gain = case
    g = np.load(os.path.join(GOLDEN, 'upfirdn2d.npz'))
    x = t(g[f'{name}_x'])
    f = t(g[f'{name}_f']) if g[f'{name}_f'].size else None
    y = upfirdn2d.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain, impl='ref')
    np.testing.assert_allclose(y.numpy(), g[f'{name}_y'], rtol=1e-5, atol=2e-6)


def test_upfirdn2d_helpers_match_reference_golden():
    g = np.load(os.path.join(GOLDEN, 'upfirdn2d.npz'))
    x, f = t(g['wrap_x']), t(g['wrap_f'])
    np.testing.assert_allclose(upfirdn2d.filter2d(x, f, impl='ref').numpy(), g['wrap_filter2d'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(upfirdn2d.upsample2d(x, f, impl='ref').numpy(), g['wrap_upsample2d'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(upfirdn2d.downsample2d(x, f, impl='ref').numpy(), g['wrap_downsample2d'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(upfirdn2d.setup_filter([1, 3, 3, 1]).numpy(), g['setup_1331'], rtol=1e-7)
    np.testing.assert_allclose(upfirdn2d.setup_filter([1, 2, 3], flip_filter=True, gain=3.0).numpy(), g['setup_flip_gain'], rtol=1e-6)
    assert upfirdn2d._parse_padding([1, 2]) == (1, 1, 2, 2) and upfirdn2d._parse_scaling(3) == (3, 3)
    assert upfirdn2d._get_filter_size(None) == (1, 1) and upfirdn2d._get_filter_size(torch.ones(3, 5)) == (5, 3)


@pytest.mark.parametrize('case', CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv2d_resample_cpu_path_matches_reference_golden(case):
    # on CPU tensors conv2d_gradfix routes to torch.nn.functional like the reference does (conv2d_gradfix.py:51-52),
    # but upfirdn2d has no CPU kernel by design -> exercise the decomposition with impl='ref' patched in
    name, xs, ws, up, down, pad, groups, flipw, usef = case
    g = np.load(os.path.join(GOLDEN, 'conv2d_resample.npz'))
    x, w, f = t(g[f'{name}_x']), t(g[f'{name}_w']), t(g['f'])
    orig = upfirdn2d.upfirdn2d
    upfirdn2d.upfirdn2d = lambda *a, **k: orig(*a, **{**k, 'impl': 'ref'})
    try:
        y = conv2d_resample.conv2d_resample(x, w, f=(f if usef else None), up=up, down=down, padding=pad, groups=groups,
                                            flip_weight=flipw)
    finally:
        upfirdn2d.upfirdn2d = orig
    np.testing.assert_allclose(y.numpy(), g[f'{name}_y'], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('case', MODCONV_CASES, ids=[c[0] for c in MODCONV_CASES])
def test_modulated_conv2d_cpu_path_matches_reference_golden(case):
    name, n, ic, oc, k, h, up, demod, noise_kind, flipw = case
    g = np.load(os.path.join(GOLDEN, 'modulated_conv2d.npz'))
    x, w, s, f = t(g[f'{name}_x']), t(g[f'{name}_w']), t(g[f'{name}_s']), t(g['f'])
    noise = t(g[f'{name}_noise']) if g[f'{name}_noise'].size else None
    orig = upfirdn2d.upfirdn2d
    upfirdn2d.upfirdn2d = lambda *a, **k: orig(*a, **{**k, 'impl': 'ref'})
    try:
        for fused, key in ((True, 'y_fused'), (False, 'y_split')):
            y = networks.modulated_conv2d(x.clone(), w, s, noise=noise, up=up, padding=k // 2, resample_filter=f,
                                          demodulate=demod, flip_weight=flipw, fused_modconv=fused)
            assert rel_l2(y, t(g[f'{name}_{key}'])) < 2e-6
    finally:
        upfirdn2d.upfirdn2d = orig


def test_fma_forward_backward():
    a = torch.randn(2, 3, 4, 4, dtype=torch.float64, requires_grad=True)
    b = torch.randn(2, 3, 1, 1, dtype=torch.float64, requires_grad=True)
    c = torch.randn(4, 4, dtype=torch.float64, requires_grad=True)
    assert torch.allclose(fma.fma(a, b, c), a * b + c)
    assert torch.autograd.gradcheck(fma.fma, (a, b, c))


def test_block_n_choice():
    assert conv2d_gradfix.choose_block_n(3, 10000) == 16
    assert conv2d_gradfix.choose_block_n(64, 10000) == 64
    assert conv2d_gradfix.choose_block_n(512, 10000) == 256
    assert conv2d_gradfix.choose_block_n(512, 16) <= 64       # few pixel tiles -> more column tiles to fill 148 SMs


def test_weight_operand_layout():
    wl = conv2d_gradfix.weight_layout
    assert wl(24, 20) == (24, 24, 32, 64)              # rows to the power of two, channels to whole swizzle rows
    assert wl(512, 513) == (512, 512, 512, 576)
    assert wl(32, 16, phases=4) == (32, 128, 128, 64)
    assert wl(24, 8, phases=4) == (32, 128, 128, 64)   # phases start at multiples of 16 columns
    assert wl(64, 64, phases=4) == (64, 256, 256, 64)
    assert wl(3, 64) == (3, 3, 16, 64)
    with pytest.raises(RuntimeError):                   # the packing kernel is the only implementation: CPU tensors are refused
        conv2d_gradfix.packed_plain(torch.randn(8, 8, 3, 3), True, 2, 1, 1)


def test_conv_desc_mirror_matches_c_struct_layout(tmp_path):
    import subprocess
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pgpp.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(pgpp_conv_desc),offsetof(pgpp_conv_desc,phase_stride),offsetof(pgpp_conv_desc,dcoef),'
                   'offsetof(pgpp_conv_desc,out),offsetof(pgpp_conv_desc,out_stride),offsetof(pgpp_conv_desc,accumulate),'
                   'offsetof(pgpp_conv_desc,operand_f16),offsetof(pgpp_conv_desc,stats_ws));return 0;}\n')
    exe = tmp_path / 'sz'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    c = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    D = custom_ops.ConvDesc
    assert c == [ctypes.sizeof(D), D.phase_stride.offset, D.dcoef.offset, D.out.offset, D.out_stride.offset, D.accumulate.offset,
                 D.operand_f16.offset, D.stats_ws.offset]


def test_wgrad_desc_mirror_matches_c_struct_layout(tmp_path):
    import subprocess
    src = tmp_path / 'szw.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pgpp.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(pgpp_wgrad_desc),offsetof(pgpp_wgrad_desc,n),offsetof(pgpp_wgrad_desc,cb),'
                   'offsetof(pgpp_wgrad_desc,products),offsetof(pgpp_wgrad_desc,out),offsetof(pgpp_wgrad_desc,workspace),offsetof(pgpp_wgrad_desc,out_scale));return 0;}\n')
    exe = tmp_path / 'szw'
    subprocess.run(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    c = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    D = custom_ops.WgradDesc
    assert c == [ctypes.sizeof(D), D.n.offset, D.cb.offset, D.products.offset, D.out.offset, D.workspace.offset, D.out_scale.offset]


def test_synthesis_chain_composition_path_matches_oracle_chain_on_cpu():
    """host logic of the callers (layer wiring, ws bookkeeping, gains, clamps) with impl='ref' on the CPU"""
    from oracle import ref_chain
    synthesis = importlib.import_module('pgpp_b200.training.synthesis')
    torch.manual_seed(0)
    net = synthesis.SynthesisChain(w_dim=32, img_resolution=64, channel_base=1024, channel_max=32, merge_channels=8).eval()
    for name, p in net.named_parameters():
        if name.endswith('noise_strength'):
            p.data.fill_(0.3)
        if name.endswith('bias') and 'affine' not in name:
            p.data.normal_()
    n = 2
    ws = torch.randn(n, net.num_ws, 32)
    pose = torch.randn(n, net.channels[8], 8, 8)
    cat = {'64': torch.randn(n, 8, 64, 64)}
    with torch.no_grad(), upfirdn2d_ref_on_cpu(upfirdn2d):
        img, parsing, tex = net(ws, pose, cat, fused=False, impl='ref', noise_mode='const')
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    rimg, rpars, rtex = ref_chain.synthesis_chain(sd, ws, pose, cat, img_resolution=64)
    assert img.shape == (n, 3, 64, 64) and parsing.shape == (n, 7, 64, 64) and tex.shape == (n, 3, 64, 64)
    assert rel_l2(img, rimg) < 1e-5 and rel_l2(parsing, rpars) < 1e-5 and rel_l2(tex, rtex) < 1e-5
    assert net.num_ws == 1 + 1 + 3 * 3 + 3


@pytest.mark.parametrize('ic,k,pad', [(3, 7, 3), (1, 3, 1), (6, 3, 1)])
def test_row_group_im2col_weights_reproduce_the_convolution(ic, k, pad):
    """host logic of the few-channel route: torch emulation of pgpp_pack_im2col + the dilated kh' x 1 GEMM over the packed weights"""
    g = torch.Generator().manual_seed(8)
    x = torch.randn(2, ic, 10, 18, generator=g, dtype=torch.float64)
    w = torch.randn(5, ic, k, k, generator=g)
    pw = conv2d_gradfix.packed_plain(w, True, 3, pad, pad, allow_im2col=True)
    im = pw.im2col
    assert im is not None and im['r'] * k * ic <= 64 and pw.c_pad == 64 and pw.kw == 1
    r, H, W = im['r'], 10, 18
    xp = torch.nn.functional.pad(x, [pad, k, pad + pad, k + r])            # generous zero border; row index yy = y' + pad
    packed = torch.zeros(2, H + pad, W, 64, dtype=torch.float64)
    for yy in range(H + pad):
        for ry in range(r):
            for kx in range(k):
                # x[n, c, yy - pad + ry, xx + kx - pad]
                packed[:, yy, :, (ry * k + kx) * ic:(ry * k + kx + 1) * ic] = xp[:, :, yy + ry + pad, kx:kx + W].permute(0, 2, 1)
    wt = pw.data.double().sum(0)[:, :5, :]                                  # [groups, o, 64]
    out = torch.zeros(2, 5, H, W, dtype=torch.float64)
    pk = torch.nn.functional.pad(packed, [0, 0, 0, 0, 0, 8])                # rows past H + pad read as zero (TMA OOB fill)
    for t_ in range(pw.kh):
        out += torch.einsum('nyxc,oc->noyx', pk[:, t_ * r:t_ * r + H], wt[t_])
    want = torch.nn.functional.conv2d(x, w.double(), padding=pad)
    assert rel_l2(out, want) < 3e-7


def test_c_abi_argument_validation_needs_no_gpu():
    """every entry point validates its arguments before the first CUDA call: error code + message through pgpp_last_error(),
    exercised here on the CPU box with dummy (never dereferenced) pointers"""
    lib = custom_ops.load_library()
    err = lambda: lib.pgpp_last_error().decode()
    p = ctypes.c_void_p(0x1000)             # non-NULL, 16-byte aligned, never touched: validation fails first
    d = custom_ops.ConvDesc()
    assert lib.pgpp_conv2d_igemm(None, None) != 0 and 'NULL' in err()
    d.act = d.wgt = d.out = 0x1000
    d.n = d.h = d.w = d.conv_h = d.conv_w = d.out_h = d.out_w = 8
    d.c_pad = 64; d.kh = d.kw = 3; d.stride = 1; d.phases = 1; d.o = d.phase_stride = 64; d.o_rows = 64; d.block_n = 48; d.products = 1
    d.a_parts = d.b_parts = 1; d.act_fn = 1
    assert lib.pgpp_conv2d_igemm(ctypes.byref(d), None) != 0 and 'block_n' in err()
    d.block_n = 64; d.products = 2
    assert lib.pgpp_conv2d_igemm(ctypes.byref(d), None) != 0 and 'products' in err()
    d.products = 3
    assert lib.pgpp_conv2d_igemm(ctypes.byref(d), None) != 0 and 'parts' in err()
    d.products = 1; d.stride = 3
    assert lib.pgpp_conv2d_igemm(ctypes.byref(d), None) != 0 and 'stride' in err()
    d.stride = 1; d.act_fn = 12
    assert lib.pgpp_conv2d_igemm(ctypes.byref(d), None) != 0 and 'no CUDA kernel found' in err()
    w = custom_ops.WgradDesc()
    w.small = w.large = w.out = w.workspace = 0x1000
    w.s_parts = w.l_parts = 1; w.n = 1; w.ca = w.cb = 8; w.ca_pad = w.cb_pad = 64; w.hs = w.ws = w.hl = w.wl = 8
    w.kh = w.kw = 3; w.pad_y = w.pad_x = 1; w.stride = 4; w.products = 1
    assert lib.pgpp_conv2d_wgrad(ctypes.byref(w), None) != 0 and 'stride' in err()
    w.stride = 1; w.cb_pad = 40
    assert lib.pgpp_conv2d_wgrad(ctypes.byref(w), None) != 0 and 'cb_pad' in err()
    assert lib.pgpp_u8_to_f32(p, 1, 4, 16, p, 4, 2, 1, None, None) != 0 and 'channel slice' in err()
    assert lib.pgpp_image_to_u8(p, 1, 0, 16, p, 0, None) != 0 and 'bad tensor size' in err()
    assert lib.pgpp_grid_sample_2d(p, p, p, 1, 1, 0, 4, 2, 2, None) != 0 and 'bad grid_sample sizes' in err()
    assert lib.pgpp_grid_sample_2d_backward(p, p, p, None, None, 1, 1, 4, 4, 2, 2, None) != 0 and 'at least one' in err()
    assert lib.pgpp_conv2d_direct(p, p, None, 1, 2, 8, 8, 8, 3, 3, 1, 1, 1.0, 1, 0.0, 1.0, -1.0, p, None, 0, 0, 0, None) != 0 and 'C * kh * kw <= 16' in err()
    assert lib.pgpp_conv2d_direct(p, p, None, 1, 1, 8, 8, 8, 3, 3, 1, 1, 1.0, 4, 0.0, 1.0, -1.0, p, None, 0, 0, 0, None) != 0 and 'linear, relu or lrelu' in err()
    taps = (ctypes.c_float * 25)(*([0.04] * 25))
    size = custom_ops.c_i64x4(1, 8, 8, 8); stride = custom_ops.c_i64x4(512, 64, 8, 1)
    assert lib.pgpp_fir_pack(p, size, stride, taps, 5, 5, 2, 2, 2, 2, 0, 1.0, p, 64, 2, None) != 0 and 'filter up to 4 x 4' in err()
    assert lib.pgpp_mix_pack(p, p, p, p, p, None, None, None, p, 1, 8, 8, 8, 64, 2, None) != 0 and 'second term' in err()
    # gradient of a fused bias_act straight into the operand format: activation, dtype and layout checks; tile count is pure host arithmetic
    assert lib.pgpp_pack_act_gradient(p, p, size, stride, 0, 4, 0.0, 1.0, -1.0, p, 64, 2, 0, None, None) != 0 and 'linear, relu, lrelu' in err()
    assert lib.pgpp_pack_act_gradient(p, p, size, stride, 3, 3, 0.2, 1.0, -1.0, p, 64, 2, 0, None, None) != 0 and 'f32, f16 or bf16' in err()
    cl = custom_ops.c_i64x4(512, 1, 64, 8)
    assert lib.pgpp_pack_act_gradient(p, p, size, cl, 0, 3, 0.2, 1.0, -1.0, p, 64, 2, 0, None, None) != 0 and 'pixel-contiguous' in err()
    assert lib.pgpp_pack_act_gradient(p, None, size, stride, 0, 3, 0.2, 1.0, -1.0, p, 64, 2, 0, None, None) != 0 and 'device pointers' in err()
    assert lib.pgpp_pack_act_gradient_tiles(512, 512) == 2048 and lib.pgpp_pack_act_gradient_tiles(513, 513) == 5 * 513
    assert lib.pgpp_pack_act_gradient_tiles(9, 7) == 2 and lib.pgpp_pack_act_gradient_tiles(40, 36) == 1 * 20


@pytest.mark.parametrize('k,pad,opad', [(1, 0, 0), (1, 0, 1), (2, 0, 0), (2, 1, 1), (3, 0, 0), (3, 1, 1), (3, 2, 1), (4, 1, 0), (5, 2, 1), (5, 4, 0), (7, 3, 1)])
def test_stride2_transposed_convolution_phase_plan(k, pad, opad):
    """the arithmetic behind conv2d_gradfix.TCONV_PHASES (sub-kernels, paddings, output phases) executed with library convolutions on the
    CPU: equal to F.conv_transpose2d(stride=2) for every kernel size / padding / output padding the GPU path accepts"""
    import torch.nn.functional as F
    cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
    g = torch.Generator().manual_seed(500 + k)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(5, 4, k, k + (1 if k in (2, 4) else 0), generator=g, dtype=torch.float64)      # [I, O, kh, kw], also non-square
    kh, kw = w.shape[2:]
    want = F.conv_transpose2d(x, w, stride=2, padding=pad, output_padding=opad)
    y = torch.full_like(want, float('nan'))
    for ph in cg.tconv_stride2_phase_plan(kh, kw, (pad, pad)):
        view = y[:, :, ph['ry']::2, ph['rx']::2]
        if view.numel() == 0:
            continue
        if ph['ty'] == 0 or ph['tx'] == 0:
            view.zero_()
            continue
        sub = w[:, :, ph['t0y']::2, ph['t0x']::2]
        assert tuple(sub.shape[2:]) == (ph['ty'], ph['tx'])
        kern = sub.transpose(0, 1).flip([2, 3])                                     # correlation kernel [O, I, ty, tx]
        # the kernel's view of a padding p (negative = crop, rows past the input = zeros): out[q] = sum_m kern[m] * x[q + m - p]
        py, px, big = ph['pad_y'], ph['pad_x'], 12
        full = F.conv2d(F.pad(x, [big] * 4), kern)                                 # full[i] = sum_m kern[m] * x[i + m - big]
        view.copy_(full[:, :, big - py: big - py + view.shape[2], big - px: big - px + view.shape[3]])
    assert not torch.isnan(y).any() and torch.allclose(y, want, atol=1e-12), float((y - want).abs().max())
