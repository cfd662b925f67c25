"""packing pass (pgpp_pack_activations) at the training shapes: ms and algorithmic GB/s.  Usage: python tools/pack_bench.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
cg._init()
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for dt, c, res in ((torch.float32, 64, 512), (torch.float32, 128, 256), (torch.float32, 512, 64), (torch.float16, 64, 512), (torch.float16, 128, 256), (torch.float32, 64, 513), (torch.float16, 64, 513)):
    x = torch.randn(8, c, res, res, device='cuda').to(dt)
    prec = cg.precision_for(dt)
    parts = cg._PRODUCTS[prec][1]
    t = timed(lambda: cg.pack_operand(x, prec))
    nbytes = x.numel() * x.element_size() + x.numel() * 2 * (1 if prec == 'f16' else parts)
    print(f'{str(dt):14s} {c:4d} ch {res}x{res} n8 {prec}: {t:.3f} ms  {nbytes / t / 1e6:.0f} GB/s')
