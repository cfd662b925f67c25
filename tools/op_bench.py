"""Op microbench sweep of BASELINE.json configs[3] as a library: `run_ops()` returns the `ops` object of the bench.py JSON line
(and `python tools/op_bench.py` prints it).  Every distinct shape of the 512 px generator at N = 32:

  modulated_conv2d  drop-in API (fp32 NCHW in / out: demodulation + packing + tcgen05 implicit GEMM) and the hand-over route the
                    generator runs (operand format in / out, modulation folded into per-sample weights, bias + activation in the
                    epilogue), fp32-parity (bf16x2) and bf16 modes: ms, algorithmic TFLOP/s, fraction of the roofline TIME bound
                    max(FLOPs / peak, min_bytes / BW) of SURVEY 8d
  upfirdn2d         fp32 NCHW through the drop-in API and the operand-format FIR (pgpp_fir_packed): ms, GB/s, fraction of HBM copy
  bias_act          forward and backward (dx): ms, GB/s, fraction of HBM copy
  conv2d_gradfix    BASELINE configs[4] shapes (discriminator blocks, batch 8 per GPU): forward + backward, and the R1 double backward

CUDA events on the current stream; the L2 is flushed (a 256 MB buffer is rewritten) before every timed launch; median of `iters`.
Peaks: MEASURED_PEAKS.json (burst figures: each op is timed alone)."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODCONV = [  # I, O, k, H_in, up, demod   (every distinct modulated_conv2d call of one forward)
    (512, 512, 3, 8, 1, True), (512, 3, 1, 8, 1, False), (512, 512, 3, 8, 2, True), (512, 512, 3, 16, 1, True), (512, 512, 3, 16, 2, True),
    (512, 512, 3, 32, 1, True), (512, 512, 3, 32, 2, True), (512, 512, 3, 64, 1, True), (512, 3, 1, 64, 1, False), (512, 256, 3, 64, 2, True),
    (256, 256, 3, 128, 1, True), (256, 3, 1, 128, 1, False), (256, 128, 3, 128, 2, True), (128, 128, 3, 256, 1, True), (128, 3, 1, 256, 1, False),
    (128, 64, 3, 256, 2, True), (64, 64, 3, 512, 1, True), (64, 7, 1, 512, 1, False), (64, 3, 1, 512, 1, False)]
UPFIRDN = [  # C, H_in, up, down, pad, gain   
    (64, 513, 1, 1, [1, 1, 1, 1], 4), (128, 257, 1, 1, [1, 1, 1, 1], 4), (256, 129, 1, 1, [1, 1, 1, 1], 4), (512, 65, 1, 1, [1, 1, 1, 1], 4),
    (64, 512, 1, 1, [2, 2, 2, 2], 1), (128, 256, 1, 1, [2, 2, 2, 2], 1), (256, 128, 1, 1, [2, 2, 2, 2], 1),
    (64, 512, 1, 2, [1, 1, 1, 1], 1), (3, 256, 2, 1, [2, 1, 2, 1], 4), (3, 64, 2, 1, [2, 1, 2, 1], 4)]
BIASACT = [  # C, H, act, gain, clamp   
    (64, 512, 'lrelu', 2 ** 0.5, 256.0), (64, 512, 'linear', 1.0, None), (64, 512, 'relu', 2 ** 0.5, None), (128, 256, 'lrelu', 2 ** 0.5, 256.0),
    (256, 128, 'lrelu', 2 ** 0.5, 256.0), (512, 64, 'lrelu', 2 ** 0.5, 256.0), (3, 512, 'linear', 1.0, 256.0), (7, 512, 'linear', 1.0, 256.0)]
TRAINCONV = [(8, 64, 64, 512), (8, 64, 128, 256), (8, 128, 256, 128), (8, 256, 512, 64), (8, 512, 512, 32)]    # n, I, O, res (D blocks, cfg 5)


def _mods():
    from __graft_entry__ import load_pkg
    load_pkg()
    names = ('bias_act', 'upfirdn2d', 'conv2d_gradfix')
    ba, up, cg = (importlib.import_module(f'pgpp_b200.torch_utils.ops.{m}') for m in names)
    return ba, up, cg, importlib.import_module('pgpp_b200.training.networks')


def run_ops(device='cuda:0', n=32, iters=4, train_shapes=True):
    ba, up, cg, nets = _mods()
    pkp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    pk = json.load(open(pkp)) if os.path.isfile(pkp) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
    BW, PEAK = pk['hbm_gbs'] * 1e9, pk['bf16_tflops'] * 1e12
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    old_prec = cg.fp32_precision

    def timeit(fn, warm=1):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2] * 1e-3

    f = up.setup_filter([1, 3, 3, 1]).to(device)
    r3 = lambda v: round(float(v), 4)
    ops = {'n': n, 'peaks': {'hbm_gbs': pk['hbm_gbs'], 'bf16_tflops_burst': pk['bf16_tflops'], 'source': 'MEASURED_PEAKS.json' if os.path.isfile(pkp) else 'fallback'},
           'timing': f'CUDA events, L2 flushed before each launch, median of {iters}',
           'modulated_conv2d': [], 'upfirdn2d': [], 'bias_act': [], 'conv2d_gradfix_train': []}
    with torch.no_grad():
        for (ic, oc, k, h, upf, demod) in MODCONV:
            torch.manual_seed(0)
            x = torch.randn(n, ic, h, h, device=device); w = torch.randn(oc, ic, k, k, device=device); s = torch.randn(n, ic, device=device) * 0.5 + 1
            nz = torch.randn(h * upf, h * upf, device=device) * 0.1 if demod else None
            b = torch.randn(oc, device=device)
            flops = 2.0 * n * oc * ic * k * k * h * h
            min_bytes = 4.0 * (n * ic * h * h + n * oc * (h * upf) ** 2) + 4.0 * oc * ic * k * k + 4 * n * ic
            row = {'shape': f'{ic}->{oc} k{k} {h}->{h * upf}', 'gflop': r3(flops / 1e9), 'ai_fp32': r3(flops / min_bytes)}
            kw = dict(noise=nz, up=upf, padding=k // 2, resample_filter=f, demodulate=demod, flip_weight=(upf == 1))
            packed_in = ic % 64 == 0 and h * h >= 128
            packed_out = oc % 16 == 0 and h * upf >= 16
            for prec in ('bf16x2', 'bf16'):
                cg.fp32_precision = prec
                parts = cg._PRODUCTS[prec][1]
                bound = max(flops / (PEAK / (3 if prec == 'bf16x2' else 1)), min_bytes / BW)      # the fp32-parity mode issues 3 MMA products per FLOP
                t_api = timeit(lambda: nets.modulated_conv2d(x, w, s, **kw))
                xin = cg.pack_operand(x, prec) if packed_in else x
                outp = cg.PackedAct(cg.PackedAct.empty(n, h * upf, h * upf, oc, parts, device), oc) if packed_out else None
                ekw = dict(bias=b, act='lrelu', gain=2 ** 0.5, clamp=256.0) if demod else dict(bias=b, act='linear', clamp=256.0)
                t_fused = timeit(lambda: nets.modulated_conv2d_fused_act(xin, w, s, noise=nz, up=upf, padding=k // 2, resample_filter=f, demodulate=demod,
                                                                        flip_weight=(upf == 1), out_packed=outp, **ekw))
                row[prec] = {'api_ms': r3(t_api * 1e3), 'api_tflops': r3(flops / t_api / 1e12), 'fused_ms': r3(t_fused * 1e3),
                             'fused_tflops': r3(flops / t_fused / 1e12), 'fused_frac_of_peak': r3(flops / t_fused / PEAK),
                             'bound_ms': r3(bound * 1e3), 'fused_frac_of_bound': r3(bound / t_fused)}
                del xin, outp
            ops['modulated_conv2d'].append(row)
            del x
        cg.fp32_precision = old_prec
        for (c, h, u, dn, pad, gain) in UPFIRDN:
            x = torch.randn(n, c, h, h, device=device)
            fn = lambda: up.upfirdn2d(x, f, up=u, down=dn, padding=pad, gain=gain, impl='cuda')
            y = fn()
            t = timeit(fn)
            byt = 4.0 * (x.numel() + y.numel())
            row = {'x': [n, c, h, h], 'up': u, 'down': dn, 'padding': pad, 'ms': r3(t * 1e3), 'gbs': r3(byt / t / 1e9), 'frac_of_hbm': r3(byt / t / BW)}
            if u == 1 and c % 8 == 0:       # the same FIR on the operand format (what the generator's encoder chains run)
                xp = cg.pack_operand(x, 'bf16x2')
                outp = cg.fir_packed(xp, f, down=dn, padding=pad, gain=gain)
                tp = timeit(lambda: cg.fir_packed(xp, f, down=dn, padding=pad, gain=gain, out=outp))
                bp = 2.0 * (xp.data.numel() + outp.data.numel())
                row['operand_format'] = {'ms': r3(tp * 1e3), 'gbs': r3(bp / tp / 1e9), 'frac_of_hbm': r3(bp / tp / BW)}
                del xp, outp
            ops['upfirdn2d'].append(row)
            del x, y
    for (c, h, act, gain, clamp) in BIASACT:
        x = torch.randn(n, c, h, h, device=device); b = torch.randn(c, device=device)
        with torch.no_grad():
            t = timeit(lambda: ba.bias_act(x, b, act=act, gain=gain, clamp=clamp, impl='cuda'))
        byt = 4.0 * (2 * x.numel() + c)
        row = {'x': [n, c, h, h], 'act': act, 'gain': r3(gain), 'clamp': clamp, 'fwd_ms': r3(t * 1e3), 'fwd_gbs': r3(byt / t / 1e9), 'fwd_frac_of_hbm': r3(byt / t / BW)}
        if act != 'linear' or clamp is not None:
            xr = x.detach().requires_grad_(True)
            y = ba.bias_act(xr, b, act=act, gain=gain, clamp=clamp, impl='cuda')
            dy = torch.randn_like(y)
            tb = timeit(lambda: torch.autograd.grad(y, xr, dy, retain_graph=True))
            bb = 4.0 * 3 * x.numel()
            row.update({'bwd_ms': r3(tb * 1e3), 'bwd_gbs': r3(bb / tb / 1e9), 'bwd_frac_of_hbm': r3(bb / tb / BW)})
            del xr, y, dy
        ops['bias_act'].append(row)
        del x
    if train_shapes:
        def step(x, w, r1):
            x = x.detach().requires_grad_(True)
            y = cg.conv2d(x, w, padding=1)
            if r1:      # loss_fullbody.py:264-274: gradient penalty, second-order graph through the convolution
                gx, = torch.autograd.grad(y.sum(), [x], create_graph=True)
                gx.square().sum().backward()
            else:
                y.square().sum().backward()
            w.grad = None
        for (bn, c, o, res) in TRAINCONV:
            x = torch.randn(bn, c, res, res, device=device)
            w = (torch.randn(o, c, 3, 3, device=device) * 0.05).requires_grad_(True)
            flops = 2.0 * bn * res * res * c * o * 9
            t1 = timeit(lambda: step(x, w, False)); t2 = timeit(lambda: step(x, w, True))
            ops['conv2d_gradfix_train'].append({'shape': f'{c}->{o} k3 {res}x{res} n{bn}', 'fwd_bwd_ms': r3(t1 * 1e3), 'fwd_bwd_tflops': r3(3 * flops / t1 / 1e12),
                                                'fwd_bwd_r1_ms': r3(t2 * 1e3), 'fwd_bwd_r1_tflops': r3(5 * flops / t2 / 1e12)})
            del x, w
    cg.fp32_precision = old_prec
    del flush
    torch.cuda.empty_cache()
    return ops


if __name__ == '__main__':
    print(json.dumps(run_ops(), indent=1))
