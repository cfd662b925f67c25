"""Op microbench sweep (BASELINE.json configs[3]): modulated_conv2d, upfirdn2d and bias_act at every distinct shape of the
512 px generator's synthesis path, batch 1 and 32, fp32 tensors through the drop-in API.
Each line: CUDA-event time, achieved TFLOP/s or GB/s on ALGORITHMIC work, fraction of the roofline bound
max(FLOPs/peak, min_bytes/BW) and parity (relative L2) against the PyTorch ref path (impl='ref' / library conv) on the same device.
    python tools/op_sweep.py [out.md]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
ba = importlib.import_module('pgpp_b200.torch_utils.ops.bias_act')
up = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
nets = importlib.import_module('pgpp_b200.training.networks')
DEV = 'cuda:0'
pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
BW, PEAK = pk['hbm_gbs'] * 1e9, pk['bf16_tflops'] * 1e12      # burst figures: kernels are timed alone


def timeit(fn, iters=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


MODCONV = [  # I, O, k, H_in, up, demod   
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
    (256, 128, 'lrelu', 2 ** 0.5, 256.0), (512, 64, 'lrelu', 2 ** 0.5, 256.0), (512, 16, 'linear', 1.0, None), (3, 512, 'linear', 1.0, 256.0),
    (7, 512, 'linear', 1.0, 256.0)]


def main():
    out = []
    torch.backends.cudnn.allow_tf32 = False
    f = up.setup_filter([1, 3, 3, 1]).to(DEV)
    out.append(f'# Op sweep on {torch.cuda.get_device_name(0)} (peaks: HBM {BW / 1e9:.0f} GB/s, bf16 {PEAK / 1e12:.0f} TFLOP/s burst, MEASURED_PEAKS.json)\n')
    out.append('## modulated_conv2d (fp32 NCHW in/out through the drop-in API: demod + pack + igemm launches; precision bf16x2 unless noted)\n')
    out.append('| I->O k | H_in->H_out | N | ms | TFLOP/s (alg.) | roofline bound ms | frac | rel-L2 vs lib fp32 | bf16 ms | bf16 TFLOP/s | bf16 frac |')
    out.append('|---|---|---|---|---|---|---|---|---|---|---|')
    with torch.no_grad():
        for (ic, oc, k, h, upf, demod) in MODCONV:
            for n in (1, 32):
                torch.manual_seed(0)
                x = torch.randn(n, ic, h, h, device=DEV); w = torch.randn(oc, ic, k, k, device=DEV); s = torch.randn(n, ic, device=DEV) * 0.5 + 1
                nz = torch.randn(h * upf, h * upf, device=DEV) * 0.1 if demod else None
                kw = dict(noise=nz, up=upf, padding=k // 2, resample_filter=f, demodulate=demod, flip_weight=(upf == 1))
                flops = 2.0 * n * oc * ic * k * k * h * h
                min_bytes = 4.0 * (n * ic * h * h + n * oc * (h * upf) ** 2) + 4.0 * oc * ic * k * k + 4 * n * ic
                bound = max(flops / PEAK, min_bytes / BW)
                res = {}
                for prec in ('bf16x2', 'bf16'):
                    cg.fp32_precision = prec
                    t = timeit(lambda: nets.modulated_conv2d(x, w, s, **kw))
                    res[prec] = t
                cg.fp32_precision = 'bf16x2'
                got = nets.modulated_conv2d(x, w, s, **kw)
                cg.enabled = False      # library (cuDNN) fp32 path of the same function = the reference composition
                want = nets.modulated_conv2d(x, w, s, **{**kw, 'noise': nz}) if False else None
                cg.enabled = True
                # reference: non-fused formulation with library ops
                xs = x * s.reshape(n, ic, 1, 1)
                wt = w if upf == 1 else w
                if upf == 1:
                    y = torch.nn.functional.conv2d(xs, w, padding=k // 2)
                else:
                    y = torch.nn.functional.conv_transpose2d(xs, w.transpose(0, 1), stride=2)
                    y = up.upfirdn2d(y, f, padding=[1, 1, 1, 1], gain=4, impl='ref')
                if demod:
                    d = (w.square().sum([2, 3])[None] * s.square()[:, None, :]).sum(2).add(1e-8).rsqrt()
                    y = y * d.reshape(n, oc, 1, 1) + nz
                err = rel(got, y)
                t = res['bf16x2']; tb = res['bf16']
                out.append(f'| {ic}->{oc} k{k} | {h}->{h * upf} | {n} | {t * 1e3:.3f} | {flops / t / 1e12:.1f} | {bound * 1e3:.3f} | {bound / t:.2f} | {err:.1e} | '
                           f'{tb * 1e3:.3f} | {flops / tb / 1e12:.1f} | {bound / tb:.2f} |')
                del x, got, y
        out.append('\n## upfirdn2d (fp32 NCHW, 4x4 filter from [1,3,3,1])\n')
        out.append('| C | H_in->H_out | up | down | N | ms | GB/s | frac of HBM | ref-path ms | speed-up vs ref path | max-abs vs ref path |')
        out.append('|---|---|---|---|---|---|---|---|---|---|---|')
        for (c, h, u, dn, pad, gain) in UPFIRDN:
            for n in (1, 32):
                x = torch.randn(n, c, h, h, device=DEV)
                fn = lambda impl='cuda': up.upfirdn2d(x, f, up=u, down=dn, padding=pad, gain=gain, impl=impl)
                y = fn()
                t = timeit(fn); tr = timeit(lambda: fn('ref'), iters=3, warm=1)
                byt = 4.0 * (x.numel() + y.numel())
                err = float((y - fn('ref')).abs().max())
                out.append(f'| {c} | {h}->{y.shape[2]} | {u} | {dn} | {n} | {t * 1e3:.3f} | {byt / t / 1e9:.0f} | {byt / t / BW:.2f} | {tr * 1e3:.3f} | {tr / t:.1f}x | {err:.1e} |')
                del x, y
        out.append('\n## bias_act (fp32 NCHW, bias on dim 1)\n')
        out.append('| shape | act | N | ms | GB/s | frac of HBM | ref-path ms | speed-up vs ref path | max-abs vs ref path |')
        out.append('|---|---|---|---|---|---|---|---|---|')
        for (c, h, act, gain, clamp) in BIASACT:
            for n in (1, 32):
                x = torch.randn(n, c, h, h, device=DEV); b = torch.randn(c, device=DEV)
                fn = lambda impl='cuda': ba.bias_act(x, b, act=act, gain=gain, clamp=clamp, impl=impl)
                y = fn()
                t = timeit(fn); tr = timeit(lambda: fn('ref'), iters=3, warm=1)
                byt = 4.0 * (2 * x.numel() + c)
                err = float((y - fn('ref')).abs().max())
                out.append(f'| [{n},{c},{h},{h}] | {act} g={gain:.3f} c={clamp} | {n} | {t * 1e3:.3f} | {byt / t / 1e9:.0f} | {byt / t / BW:.2f} | {tr * 1e3:.3f} | {tr / t:.1f}x | {err:.1e} |')
                del x, y
    text = '\n'.join(out) + '\n'
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'op_sweep.md')
    os.makedirs(os.path.dirname(path), exist_ok=True)
    open(path, 'w').write(text)
    print(text)


if __name__ == '__main__':
    main()
