"""GPU baseline of the REFERENCE's own kernels, same box, same shapes (VERDICT r1 item 6 / SURVEY 2.2):
  * bias_act / upfirdn2d: the reference's CUDA plugins recompiled for sm_100a (oracle/build_ref_kernels.py -> oracle/_ref/*.so, built
    from /root/reference/torch_utils/ops/{bias_act,upfirdn2d}.{cpp,cu}, called exactly as bias_act.py:153 / upfirdn2d.py:237 call them)
  * modulated_conv2d: the reference's fused formulation (networks.py:85-93: per-sample weights, cuDNN grouped conv with groups=N),
    in float32 and with TF32 allowed, plus its up=2 form (grouped conv_transpose2d, then the reference's upfirdn2d blur)
against this repo's kernels through the drop-in API.  Shapes: SURVEY Appendix A / C / D at N = 32.
    python tools/ref_kernels_baseline.py [out.md]"""
import importlib, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
from oracle import build_ref_kernels
load_pkg()
ba = importlib.import_module('pgpp_b200.torch_utils.ops.bias_act')
up = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
nets = importlib.import_module('pgpp_b200.training.networks')
DEV = 'cuda:0'
pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
BW, PEAK = pk['hbm_gbs'] * 1e9, pk['bf16_tflops'] * 1e12
N = 32
flush = None


def timeit(fn, iters=6, warm=2):
    """median of per-iteration CUDA-event times, L2 flushed (a buffer larger than the 126 MB L2 is rewritten) before each"""
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
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


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    rb = build_ref_kernels.load('ref_bias_act_plugin')
    ru = build_ref_kernels.load('ref_upfirdn2d_plugin')
    out = [f'# Reference kernels recompiled for sm_100a vs this repo, {torch.cuda.get_device_name(0)}, N = {N} '
           f'(HBM copy {BW / 1e9:.0f} GB/s, bf16 burst {PEAK / 1e12:.0f} TFLOP/s from MEASURED_PEAKS.json; CUDA events, L2 flushed, median of 6)\n']
    null = torch.empty([0], device=DEV)
    act_idx = {'linear': 1, 'relu': 2, 'lrelu': 3}
    out.append('## bias_act (fp32 NCHW): reference plugin `bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)` (bias_act.cpp:32)\n')
    out.append('| shape | act | pass | reference ms | GB/s | ours ms | GB/s | frac of HBM (ours) | ours / reference | max-abs diff |')
    out.append('|---|---|---|---|---|---|---|---|---|---|')
    for (c, h, act, gain, clamp, bias) in [(64, 512, 'lrelu', 2 ** 0.5, 256.0, True), (64, 512, 'linear', 1.0, -1.0, True), (64, 512, 'relu', 2 ** 0.5, -1.0, False),
                                           (128, 256, 'lrelu', 2 ** 0.5, 256.0, True), (256, 128, 'lrelu', 2 ** 0.5, 256.0, True), (512, 64, 'lrelu', 2 ** 0.5, 256.0, True),
                                           (3, 512, 'linear', 1.0, 256.0, True)]:
        x = torch.randn(N, c, h, h, device=DEV); b = torch.randn(c, device=DEV) if bias else None
        alpha = 0.2 if act == 'lrelu' else 0.0
        ref_fn = lambda: rb.bias_act(x, b if bias else null, null, null, null, 0, 1, act_idx[act], alpha, gain, clamp)
        our_fn = lambda: ba.bias_act(x, b, act=act, gain=gain, clamp=None if clamp < 0 else clamp, impl='cuda')
        tr, to = timeit(ref_fn), timeit(our_fn)
        byt = 4.0 * (2 * x.numel() + c)
        out.append(f'| [{N},{c},{h},{h}] | {act} g={gain:.3f} c={clamp} {"+b" if bias else "no-b"} | fwd | {tr * 1e3:.3f} | {byt / tr / 1e9:.0f} | {to * 1e3:.3f} | {byt / to / 1e9:.0f} | '
                   f'{byt / to / BW:.2f} | {tr / to:.2f}x | {float((ref_fn() - our_fn()).abs().max()):.1e} |')
        if act == 'lrelu' and c <= 128:
            # backward (grad = 1): dx from dy and the saved output, as BiasActCudaGrad.forward calls it (bias_act.py:182)
            y = our_fn(); dy = torch.randn_like(y)
            ref_b = lambda: rb.bias_act(dy, null, null, y, null, 1, 1, act_idx[act], alpha, gain, clamp)
            xr = x.detach().requires_grad_(True)
            yo = ba.bias_act(xr, b, act=act, gain=gain, clamp=clamp, impl='cuda')
            our_b = lambda: torch.autograd.grad(yo, xr, dy, retain_graph=True)[0]
            tr, to = timeit(ref_b), timeit(our_b)
            byt = 4.0 * 3 * x.numel()
            out.append(f'| [{N},{c},{h},{h}] | {act} | bwd dx | {tr * 1e3:.3f} | {byt / tr / 1e9:.0f} | {to * 1e3:.3f} | {byt / to / 1e9:.0f} | {byt / to / BW:.2f} | {tr / to:.2f}x | '
                       f'{float((ref_b() - our_b()).abs().max()):.1e} |')
            del y, dy, xr, yo
        del x
    f = up.setup_filter([1, 3, 3, 1]).to(DEV)
    out.append('\n## upfirdn2d (fp32 NCHW, 4x4 filter): reference plugin `upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain)` (upfirdn2d.cpp:16)\n')
    out.append('| x | up | down | padding | reference ms | GB/s | ours ms | GB/s | frac of HBM (ours) | ours / reference | max-abs diff |')
    out.append('|---|---|---|---|---|---|---|---|---|---|---|')
    for (c, h, u, d, pad, gain) in [(64, 513, 1, 1, [1, 1, 1, 1], 4.0), (128, 257, 1, 1, [1, 1, 1, 1], 4.0), (256, 129, 1, 1, [1, 1, 1, 1], 4.0), (512, 65, 1, 1, [1, 1, 1, 1], 4.0),
                                    (64, 512, 1, 1, [2, 2, 2, 2], 1.0), (128, 256, 1, 1, [2, 2, 2, 2], 1.0), (256, 128, 1, 1, [2, 2, 2, 2], 1.0),
                                    (64, 512, 1, 2, [1, 1, 1, 1], 1.0), (3, 256, 2, 1, [2, 1, 2, 1], 4.0), (3, 64, 2, 1, [2, 1, 2, 1], 4.0)]:
        x = torch.randn(N, c, h, h, device=DEV)
        ref_fn = lambda: ru.upfirdn2d(x, f, u, u, d, d, pad[0], pad[1], pad[2], pad[3], False, gain)
        our_fn = lambda: up.upfirdn2d(x, f, up=u, down=d, padding=pad, gain=gain, impl='cuda')
        y = our_fn()
        tr, to = timeit(ref_fn), timeit(our_fn)
        byt = 4.0 * (x.numel() + y.numel())
        out.append(f'| [{N},{c},{h},{h}] | {u} | {d} | {pad} | {tr * 1e3:.3f} | {byt / tr / 1e9:.0f} | {to * 1e3:.3f} | {byt / to / 1e9:.0f} | {byt / to / BW:.2f} | {tr / to:.2f}x | '
                   f'{float((ref_fn() - y).abs().max()):.1e} |')
        del x, y
    out.append('\n## modulated_conv2d: reference fused formulation (networks.py:85-93, cuDNN grouped conv, groups = N; weights [N*O, I, k, k] rebuilt per call as the '
               'reference does) vs this repo (`modulated_conv2d` drop-in, fp32 NCHW in / out, demodulation + packing + tcgen05 implicit GEMM)\n')
    out.append('| I->O k | H_in->H_out | cuDNN fp32 ms | cuDNN TF32 ms | ours bf16x2 ms | TFLOP/s (alg.) | frac of bf16 burst / 3 | ours bf16 ms | TFLOP/s | frac of bf16 burst | fp32 / ours | TF32 / ours | rel-L2 ours vs cuDNN fp32 | rel-L2 TF32 vs fp32 |')
    out.append('|---|---|---|---|---|---|---|---|---|---|---|---|---|---|')
    with torch.no_grad():
        for (ic, oc, k, h, upf) in [(64, 64, 3, 512, 1), (128, 128, 3, 256, 1), (256, 256, 3, 128, 1), (512, 512, 3, 64, 1), (512, 512, 3, 16, 1),
                                    (128, 64, 3, 256, 2), (256, 128, 3, 128, 2), (64, 3, 1, 512, 1)]:
            torch.manual_seed(0)
            demod = k == 3
            x = torch.randn(N, ic, h, h, device=DEV); w = torch.randn(oc, ic, k, k, device=DEV); s = torch.randn(N, ic, device=DEV) * 0.5 + 1
            nz = torch.randn(h * upf, h * upf, device=DEV) * 0.1 if demod else None

            def ref_fused():
                # networks.py:62-93 (fused_modconv=True), float32
                ww = w.unsqueeze(0) * s.reshape(N, 1, ic, 1, 1)
                if demod:
                    ww = ww * (ww.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt().reshape(N, oc, 1, 1, 1)
                xx = x.reshape(1, N * ic, h, h)
                if upf == 1:
                    y = torch.nn.functional.conv2d(xx, ww.reshape(N * oc, ic, k, k), padding=k // 2, groups=N)
                else:
                    wt = ww.reshape(N, oc, ic, k, k).transpose(1, 2).reshape(N * ic, oc, k, k)     # conv2d_resample.py:125-137
                    y = torch.nn.functional.conv_transpose2d(xx, wt, stride=2, groups=N)
                    y = ru.upfirdn2d(y, f, 1, 1, 1, 1, 1, 1, 1, 1, False, 4.0)
                y = y.reshape(N, oc, h * upf, h * upf)
                return y.add_(nz) if nz is not None else y
            kw = dict(noise=nz, up=upf, padding=k // 2, resample_filter=f, demodulate=demod, flip_weight=(upf == 1))
            flops = 2.0 * N * oc * ic * k * k * h * h
            torch.backends.cudnn.allow_tf32 = False
            t32 = timeit(ref_fused, iters=4, warm=1); want = ref_fused()
            torch.backends.cudnn.allow_tf32 = True
            ttf = timeit(ref_fused, iters=4, warm=1); err_tf = rel(ref_fused(), want)
            torch.backends.cudnn.allow_tf32 = False
            res = {}
            for prec in ('bf16x2', 'bf16'):
                cg.fp32_precision = prec
                res[prec] = timeit(lambda: nets.modulated_conv2d(x, w, s, **kw), iters=4, warm=1)
            cg.fp32_precision = 'bf16x2'
            err = rel(nets.modulated_conv2d(x, w, s, **kw), want)
            a, b = res['bf16x2'], res['bf16']
            out.append(f'| {ic}->{oc} k{k} | {h}->{h * upf} | {t32 * 1e3:.3f} | {ttf * 1e3:.3f} | {a * 1e3:.3f} | {flops / a / 1e12:.1f} | {flops / a / (PEAK / 3):.2f} | {b * 1e3:.3f} | '
                       f'{flops / b / 1e12:.1f} | {flops / b / PEAK:.2f} | {t32 / a:.2f}x | {ttf / a:.2f}x | {err:.1e} | {err_tf:.1e} |')
            del x, want
    text = '\n'.join(out) + '\n'
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'ref_kernels_baseline.md')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()
