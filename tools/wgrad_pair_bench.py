"""A/B of the weight-gradient kernel's tap-pair form (<= 64 channels on the accumulator-lane side) against the one-tap-per-accumulator
form (PGPP_WGRAD_NO_PAIR) on the training iteration's 64-channel layers.  Usage: python tools/wgrad_pair_bench.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import load_pkg  # noqa: E402

load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    cases = [(8, 64, 64, 512, 'bf16x2'), (8, 64, 64, 512, 'f16'), (8, 64, 128, 512, 'bf16x2'), (8, 32, 32, 512, 'bf16x2'), (8, 64, 64, 256, 'bf16x2')]
    for n, ca, cb, res, prec in cases:
        dt = torch.float16 if prec == 'f16' else torch.float32
        dy = torch.randn(n, ca, res, res, device=DEV).to(dt)
        x = torch.randn(n, cb, res, res, device=DEV).to(dt)
        sp, lp = cg.pack_operand(dy, prec), cg.pack_operand(x, prec)
        row = []
        for flag in (None, '1'):
            if flag:
                os.environ['PGPP_WGRAD_NO_PAIR'] = flag
            else:
                os.environ.pop('PGPP_WGRAD_NO_PAIR', None)
            custom_ops.refresh_env()
            ms = timed(lambda: cg.weight_gradient(sp, lp, (ca, cb, 3, 3), 1, (1, 1), False, precision=prec, out_dtype=torch.float32))
            row.append(ms)
        os.environ.pop('PGPP_WGRAD_NO_PAIR', None)
        custom_ops.refresh_env()
        flops = 2.0 * n * res * res * ca * cb * 9
        print(f'wgrad {ca}x{cb} k3 {res}x{res} n{n} {prec}: pair {row[0]:.3f} ms ({flops / row[0] / 1e9:.0f} TFLOP/s)   one tap per accumulator '
              f'{row[1]:.3f} ms ({flops / row[1] / 1e9:.0f} TFLOP/s)')


if __name__ == '__main__':
    main()
