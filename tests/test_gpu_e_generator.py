"""GPU parity of the full 512 px generator (fused kernel route) against the fixture minted from the REAL reference
GeneratorFull_v20 (tests/golden/generator.npz) -- the north-star end-to-end check: fp32 mode, max abs error <= 1e-3 relative
to the output scale, gt_parsing fixed so no discrete decision depends on rounding."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import ref_generator

pytestmark = pytest.mark.gpu
load_pkg()
gen = importlib.import_module('pgpp_b200.training.generator')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'generator.npz')
DEV = 'cuda:0'


def pooled(t):
    return torch.nn.functional.avg_pool2d(t.float().cpu(), 8)


@pytest.fixture(scope='module')
def generator_and_inputs():
    G = gen.build_generator().eval()
    ref_generator.name_seeded_init(list(G.named_parameters()) + list(G.named_buffers()))
    G = G.to(DEV).requires_grad_(False)
    inp = {k: v.to(DEV) for k, v in ref_generator.synthetic_inputs(1, seed=0).items()}
    return G, inp


def _run(G, inp, gt=True, **kw):
    with torch.no_grad():
        return G(torch.zeros(inp['c'].shape[0], 0, device=DEV), inp['c'], inp['retain'], inp['pose'], inp['denorm_upper'], inp['denorm_lower'],
                 inp['denorm_upper_mask'], inp['denorm_lower_mask'], gt_parsing=inp['gt_parsing'] if gt else None, noise_mode='const', **kw)


@pytest.mark.parametrize('prec,tol', [('bf16x2', 5e-4), ('bf16x3', 3e-4), ('bf16', 5e-2)])
def test_fused_generator_matches_reference_golden(generator_and_inputs, prec, tol):
    G, inp = generator_and_inputs
    g = np.load(GOLDEN)
    old = cg.fp32_precision
    cg.fp32_precision = prec
    try:
        before = custom_ops.launch_count()
        img, fin, pred = _run(G, inp)
        assert custom_ops.launch_count() - before >= 100           # 100 convolutions per image on the native kernels
    finally:
        cg.fp32_precision = old
    stats = g['gt_stats']
    for name, t, scale in (('img', img, stats[0]), ('finetune', fin, stats[2]), ('parsing', pred, stats[4])):
        want = torch.from_numpy(g[f'gt_{name}_pooled'])
        err = float((pooled(t) - want).norm() / want.norm())
        assert err < tol, (prec, name, err)
    if prec != 'bf16':
        crop = img[:, :, 200:232, 240:272].cpu().numpy()
        assert np.abs(crop - g['gt_img_crop']).max() <= 1e-3 * max(1.0, stats[0])
        crop = fin[:, :, 200:232, 240:272].cpu().numpy()
        assert np.abs(crop - g['gt_finetune_crop']).max() <= 1e-3 * max(1.0, stats[2])


FULLRES = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'generator_fullres.npz')
# end-to-end bars at FULL resolution (no pooling): north star max-abs <= 1e-3 of the output scale; rel-L2 of the whole image
# bf16x2 (the benchmarked mode) holds the 1e-4 bar on all three outputs (measured 6.9e-5 / 9.0e-5 / 5.0e-5).  bf16x3 is NOT tighter
# end to end (measured 1.3e-4 on img): its operands are exact to 24 bits, but it chains twice as many MMA products into the same
# fp32 TMEM accumulator, and the tensor core's accumulation error grows with the chain length - see DESIGN.md section 2.
FULL_TOL = {'bf16x2': 1.0e-4, 'bf16x3': 2.0e-4}


@pytest.mark.parametrize('prec', ['bf16x2', 'bf16x3'])
def test_fused_generator_full_resolution_against_reference_samples(generator_and_inputs, prec):
    """un-pooled comparison with the REAL reference's outputs (oracle/make_golden_generator.py -> generator_fullres.npz): every 3rd
    (parsing: 5th) pixel of all three outputs, per-pixel"""
    G, inp = generator_and_inputs
    g = np.load(FULLRES)
    old = cg.fp32_precision
    cg.fp32_precision = prec
    try:
        img, fin, pred = _run(G, inp)
    finally:
        cg.fp32_precision = old
    report = {}
    for name, t, key, step in (('img', img, 'gt_img_s3', 3), ('finetune', fin, 'gt_finetune_s3', 3), ('parsing', pred, 'gt_parsing_s5', 5)):
        want = torch.from_numpy(g[key]).double()
        got = t[:, :, ::step, ::step].cpu().double()
        rel = float((got - want).norm() / want.norm())
        mx = float((got - want).abs().max())
        scale = float(want.abs().max())
        report[name] = (rel, mx / scale)
        assert mx <= 1e-3 * max(1.0, scale), (prec, name, mx, scale)
        assert rel <= FULL_TOL[prec], (prec, name, rel)
    print(f'full-resolution parity {prec}: ' + ', '.join(f'{k} rel-L2 {v[0]:.2e} max-abs/scale {v[1]:.2e}' for k, v in report.items()))


def test_composition_route_on_gpu_and_batch_independence(generator_and_inputs):
    """drop-in composition (modulated_conv2d + bias_act + conv2d_resample calls) equals the fused route; samples are independent"""
    G, inp = generator_and_inputs
    a = _run(G, inp, fused=True)
    b = _run(G, inp, fused=False)
    for x, y in zip(a, b):
        assert float((x - y).norm() / y.norm()) < 3e-4
    inp2 = {k: torch.cat([v, v.flip(-1)]) for k, v in inp.items()}
    c = _run(G, inp2, fused=True)
    for x, y in zip(c, a):
        # not bit-identical: library reductions (instance norm, sums) pick batch-size dependent algorithms; a kernel-level
        # cross-sample leak (e.g. resident weights reloaded too early) shows up as >= 1e-3
        assert float((x[:1] - y).norm() / y.norm()) < 2e-4


def test_cuda_graph_replay_is_bit_identical(generator_and_inputs):
    G, inp = generator_and_inputs
    x = {k: v for k, v in inp.items() if k != 'gt_parsing'}
    eager = _run(G, inp, gt=False)
    gg = gen.GraphedGenerator(G, x)
    for _ in range(2):
        out = gg(x)
    torch.cuda.synchronize()
    for a, b in zip(out, eager):
        assert torch.equal(a, b)


def test_cuda_graph_gt_parsing_is_a_refreshed_static_input_and_bad_calls_raise(generator_and_inputs):
    G, inp = generator_and_inputs
    x = {k: v for k, v in inp.items() if k != 'gt_parsing'}
    gt_a = inp['gt_parsing']
    gt_b = (gt_a + 1) % 7
    gg = gen.GraphedGenerator(G, x, gt_parsing=gt_a)
    out_b = [t.clone() for t in gg(x, gt_parsing=gt_b)]
    torch.cuda.synchronize()
    want_b = _run(G, dict(x, gt_parsing=gt_b))
    for a, b in zip(out_b, want_b):
        assert torch.equal(a, b)                       # the new parsing map was used, not the one baked in at capture
    with pytest.raises(ValueError):
        gg(x)                                          # captured with gt_parsing: must be supplied
    with pytest.raises(KeyError):
        gg(dict(x, extra=x['c']))
    with pytest.raises(KeyError):
        gg({k: v for k, v in x.items() if k != 'pose'})


@pytest.mark.parametrize('down', [1, 2])
@pytest.mark.parametrize('packed_input', [False, True])
def test_resblock_hand_over_route_matches_composition(down, packed_input):
    """ResBlock (networks.py:287-316): x packed once for skip + conv0, conv0 -> conv1 in operand format, conv1's epilogue adds
    into y (relu, then += skip) - against the op-by-op composition on the CPU oracle path"""
    if packed_input and down == 2:
        pytest.skip('the down=2 block starts with a FIR pass on the NCHW tensor')
    torch.manual_seed(3)
    blk = gen.ResBlock(64, 128 if down == 2 else 64, 3, activation='relu', down=down).eval().requires_grad_(False)
    for p in blk.parameters():
        p.copy_(torch.randn_like(p))
    x = torch.randn(2, 64, 48, 80)
    with torch.no_grad():
        from helpers import upfirdn2d_ref_on_cpu
        upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
        with upfirdn2d_ref_on_cpu(upf):
            want = blk(x.double(), fused=False, impl='ref') if False else blk(x, fused=False, impl='ref')
        blk = blk.to(DEV)
        xin = x.to(DEV)
        if packed_input:
            data = cg._plugin.pack_activations(xin, None, 64, 2) if cg._init() else None
            xin = cg.PackedAct(data, 64)
        blk(xin, fused=True)                            # first call packs the weights (one launch per tensor, then cached)
        launches = custom_ops.launch_count()
        got = blk(xin, fused=True)
        n_launch = custom_ops.launch_count() - launches
    rel = ((got.cpu() - want).norm() / want.norm()).item()
    assert rel < 1e-4, rel
    # down=1: [pack] + skip + conv0 + conv1;  down=2: FIR decimation + pack + skip, blur + pack + conv0, conv1
    assert n_launch == ((3 if packed_input else 4) if down == 1 else 7), n_launch


@pytest.mark.parametrize('c,hw', [(128, (32, 48)), (64, (24, 40))])
def test_spade_epilogue_equals_the_two_pass_route(c, hw):
    """Spade_Norm_Block fused into the gamma|beta GEMM's epilogue (pgpp_conv_desc.spade_*) against the GEMM -> float32 ->
    pgpp_spade_modulate_pack route: same arithmetic on the same accumulators, so the operands must be bit-identical; and both
    against the module's op-by-op forward."""
    torch.manual_seed(5)
    blk = gen.Spade_Norm_Block(128, c).to(DEV).eval().requires_grad_(False)
    for p in blk.parameters():
        p.copy_(torch.randn_like(p))
    h, w = hw
    x = torch.randn(2, c, h, w, device=DEV) * 2 + 0.5
    feats = torch.randn(2, 128, h, w, device=DEV)
    var, mean = torch.var_mean(x, dim=(2, 3), unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    fp = cg.PackedAct(cg._plugin.pack_activations(feats, None, 128, 2), 128) if cg._init() else None
    outs = []
    for flag in (True, False):
        gen.FUSE_SPADE_EPILOGUE = flag
        try:
            outs.append(blk.fused_packed(x, mean, rstd, fp, pre_gain=1.25))
        finally:
            gen.FUSE_SPADE_EPILOGUE = True
    assert torch.equal(outs[0].data[..., :c], outs[1].data[..., :c])
    with torch.no_grad():
        want = torch.relu(blk(x, feats, fused=False)) * 1.25
    got = outs[0].to_nchw()
    assert ((got - want).norm() / want.norm()).item() < 1e-4
