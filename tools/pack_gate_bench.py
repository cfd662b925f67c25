"""pgpp_pack_act_gradient (dy * act'(y) -> operand format + per-tile channel sums, one pass) against the three passes it replaces
(bias_act gradient kernel, packing pass, pgpp_sum_hw) on the discriminator's largest fp16 / fp32 maps.  Usage: python tools/pack_gate_bench.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import load_pkg  # noqa: E402

load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
ba = importlib.import_module('pgpp_b200.torch_utils.ops.bias_act')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'


def timed(fn, iters=20):
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
    cg._init(); ba._init()
    plug = custom_ops.get_plugin('conv2d_plugin')
    fn = ba._bias_act_cuda(dim=1, act='lrelu', alpha=None, gain=None, clamp=256.0)
    spec, alpha, gain, clamp = ba._resolve('lrelu', None, None, 256.0)
    for dt, c, res in ((torch.float16, 64, 512), (torch.float16, 128, 256), (torch.float32, 64, 512), (torch.float32, 128, 256)):
        prec = cg.precision_for(dt)
        parts = cg._PRODUCTS[prec][1]
        dy = torch.randn(8, c, res, res, device=DEV).to(dt)
        y = torch.randn(8, c, res, res, device=DEV).to(dt)
        null = ba._null_tensor

        def three():
            with torch.no_grad():
                dx = fn.Grad.apply(dy, null, null, y)
                cg.pack_operand(dx, prec)
                plug.sum_hw(dx).sum(0)

        def one():
            _, sums = plug.pack_act_gradient(dy, y, 3, alpha, gain, clamp, c, parts, f16=prec == 'f16', want_sums=True)
            sums.sum([0, 2])

        t3, t1 = timed(three), timed(one)
        nbytes = dy.numel() * dy.element_size() * 2 + dy.numel() * (2 if prec == 'f16' else 2 * parts)
        print(f'{str(dt):14s} {c:4d} ch {res}x{res} n8: three passes {t3:.3f} ms   one pass {t1:.3f} ms ({nbytes / t1 / 1e6:.0f} GB/s algorithmic)')


if __name__ == '__main__':
    main()
