"""Quick per-op timing on one GPU (CUDA events, inputs larger than L2).  Not the contract bench (bench.py)."""
import importlib
import json
import os
import sys

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


def timeit(fn, iters=10, warm=3):
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


def main():
    res = []
    n = int(os.environ.get('MB_N', 8))
    with torch.no_grad():
        x = torch.randn(n, 64, 512, 512, device=DEV)
        b = torch.randn(64, device=DEV)
        for name, fn in [('bias_act lrelu f32 ours', lambda: ba.bias_act(x, b, act='lrelu', clamp=256.0)),
                         ('bias_act lrelu f32 torch-ref', lambda: ba.bias_act(x, b, act='lrelu', clamp=256.0, impl='ref')),
                         ('copy_ (roofline probe)', lambda: torch.empty_like(x).copy_(x))]:
            s = timeit(fn)
            res.append((name, s, 2 * x.numel() * 4 / s / 1e9, 'GB/s'))
        f = up.setup_filter([1, 3, 3, 1]).to(DEV)
        xo = torch.randn(n, 64, 513, 513, device=DEV)
        for name, fn, byt in [
            ('upfirdn2d blur 513->512 ours', lambda: up.upfirdn2d(xo, f, padding=[1, 1, 1, 1], gain=4), (xo.numel() + n * 64 * 512 * 512) * 4),
            ('upfirdn2d blur 513->512 torch-ref', lambda: up.upfirdn2d(xo, f, padding=[1, 1, 1, 1], gain=4, impl='ref'), (xo.numel() + n * 64 * 512 * 512) * 4),
            ('upfirdn2d blur 512->513 ours', lambda: up.upfirdn2d(x, f, padding=[2, 2, 2, 2]), (x.numel() + n * 64 * 513 * 513) * 4),
            ('upfirdn2d down2 512->256 ours', lambda: up.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]), (x.numel() + n * 64 * 256 * 256) * 4),
        ]:
            s = timeit(fn)
            res.append((name, s, byt / s / 1e9, 'GB/s'))
        del xo
        # modulated convs (fp32 API): TFLOP/s on algorithmic FLOPs
        for (ic, oc, h, upf) in [(512, 512, 32, 1), (512, 512, 64, 1), (256, 256, 128, 1), (128, 128, 256, 1), (64, 64, 512, 1),
                                 (512, 512, 32, 2), (128, 64, 256, 2)]:
            xx = torch.randn(n, ic, h, h, device=DEV)
            w = torch.randn(oc, ic, 3, 3, device=DEV)
            s_ = torch.randn(n, ic, device=DEV)
            nz = torch.randn(h * upf, h * upf, device=DEV)
            flops = 2.0 * n * oc * ic * 9 * h * h
            for prec in ('bf16', 'bf16x2', 'bf16x3'):
                cg.fp32_precision = prec
                s = timeit(lambda: nets.modulated_conv2d(xx, w, s_, noise=nz, up=upf, padding=1, resample_filter=f, flip_weight=(upf == 1)), iters=5, warm=2)
                res.append((f'modconv {ic}->{oc} k3 {h}->{h * upf} N={n} {prec} (fp32 NCHW in/out, incl. pack)', s, flops / s / 1e12, 'TFLOP/s'))
            del xx
    for name, s, v, u in res:
        print(f'{name:80s} {s * 1e3:9.3f} ms  {v:9.1f} {u}')
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'microbench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
