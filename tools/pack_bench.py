"""HBM roofline of the operand-preparation kernels (pack_nchw, spade_pack, im2col) at generator sizes (batch 32):
    python tools/pack_bench.py"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
load_pkg()
from pgpp_b200.torch_utils import custom_ops
plugin = custom_ops.get_plugin('conv2d_plugin')
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
hbm = 6549.1
dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes):
    print(f'{name:58s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s  {nbytes / ms / 1e6 / hbm * 100:5.1f}% of HBM copy', flush=True)


for (n, c, h, w, parts) in [(32, 64, 512, 320, 2), (32, 128, 256, 160, 2), (32, 256, 128, 80, 2), (32, 64, 513, 321, 2), (32, 64, 512, 320, 1)]:
    x = torch.randn(n, c, h, w, device=dev)
    ms = timeit(lambda: plugin.pack_activations(x, None, (c + 63) // 64 * 64, parts))
    report(f'pack_nchw {n}x{c}x{h}x{w} f32 -> {parts} part(s)', ms, x.numel() * (4 + 2 * parts))
xpad = torch.randn(32, 64, 513, 516, device=dev)[..., :513]
ms = timeit(lambda: plugin.pack_activations(xpad, None, 64, 2))
report('pack_nchw 32x64x513x513 (row pitch 516) f32 -> 2 parts', ms, xpad.numel() * 8)
up = custom_ops.get_plugin('upfirdn2d_plugin')
f = torch.tensor([1., 3., 3., 1.], device=dev); f = torch.outer(f, f); f = f / f.sum()
xin = torch.randn(32, 64, 512, 512, device=dev)
for align in (1, 4):
    ms = timeit(lambda: up.upfirdn2d(xin, f, 1, 1, 1, 1, 2, 2, 2, 2, False, 1.0, row_align=align))
    report(f'fir blur 32x64x512x512 -> 513x513 row_align={align}', ms, xin.numel() * 4 + 32 * 64 * 513 * 513 * 4)
for (n, c, h, w) in [(32, 128, 256, 160), (32, 64, 512, 320)]:
    x = torch.randn(n, c, h, w, device=dev)
    gb = torch.randn(n, 2 * c, h, w, device=dev)
    mean = torch.randn(n, c, device=dev); rstd = torch.rand(n, c, device=dev) + 0.5
    ms = timeit(lambda: plugin.spade_modulate_pack(x, mean, rstd, gb, c, 2, 1.0))
    report(f'spade_pack {n}x{c}x{h}x{w} -> 2 parts', ms, x.numel() * (12 + 4))
x = torch.randn(32, 3, 512, 320, device=dev)
ms = timeit(lambda: plugin.pack_im2col(x, None, 7, 3, 3, 3, 2))
report('im2col 32x3x512x320 k7 r3 -> 2 parts', ms, x.numel() * 4 + 32 * 515 * 320 * 64 * 2 * 2)

cg = __import__('importlib').import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
for (ic, oc, k, h, w, packed) in [(1, 64, 3, 512, 512, True), (1, 64, 3, 512, 512, False), (5, 64, 1, 512, 512, False), (1, 128, 3, 256, 256, True)]:
    x = torch.randn(32, ic, h, w, device=dev)
    wt = torch.randn(oc, ic, k, k, device=dev)
    op = cg.PackedAct(cg.PackedAct.empty(32, h, w, oc, 2, dev), oc) if packed else None
    ms = timeit(lambda: cg.direct_conv(x, wt, None, act='relu', out_packed=op))
    report(f'direct conv {ic}->{oc} k{k} 32x{h}x{w} -> {"2 bf16 parts" if packed else "f32 NCHW"}', ms, x.numel() * 4 + 32 * oc * h * w * 4)

f16 = [1., 3., 3., 1.]; taps = [a * b / 64. for a in f16 for b in f16]
for (n, c, h, w) in [(32, 64, 512, 512), (32, 128, 256, 256)]:
    x = torch.randn(n, c, h, w, device=dev)
    ms = timeit(lambda: plugin.fir_pack(x, taps, 4, 4, 2, 2, 2, 2, False, 1.0, c, 2))
    report(f'fir_pack {n}x{c}x{h}x{w} -> {h + 1}x{w + 1}, 2 parts', ms, x.numel() * 4 + n * c * (h + 1) * (w + 1) * 4)
