"""up = 2 synthesis layer on operand-format tensors, batch 32: per-phase transposed-conv GEMMs + blur pass (UP2_PHASES) vs the one-launch
4x-MAC polyphase GEMM:   python tools/up2_bench.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
nets = importlib.import_module('pgpp_b200.training.networks')
upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
cg.fp32_precision = sys.argv[1] if len(sys.argv) > 1 else 'bf16x2'
parts = cg._PRODUCTS[cg.fp32_precision][1]
dev = 'cuda:0'
cg._init()
f = upf.setup_filter([1, 3, 3, 1]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for ic, oc, res in [(512, 512, 32), (512, 256, 64), (256, 128, 128), (128, 64, 256)]:
    n = 32
    x = torch.randn(n, ic, res, res, device=dev); w = torch.randn(oc, ic, 3, 3, device=dev); s = torch.rand(n, ic, device=dev) + 0.5
    b = torch.randn(oc, device=dev); nz = torch.randn(2 * res, 2 * res, device=dev)
    xp = cg.PackedAct(cg._plugin.pack_activations(x, None, ic, parts), ic)
    out = cg.PackedAct(cg.PackedAct.empty(n, 2 * res, 2 * res, oc, parts, dev), oc)
    del x
    res_ms = {}
    for flag in (True, 'taps4', False):
        nets.UP2_PHASES = bool(flag)
        nets.UP2_PHASES_MIN_IO, nets.UP2_TAPS4_MIN_IO = ((1 << 30), 0) if flag == 'taps4' else (0, 1 << 30)
        def run():
            with torch.no_grad():
                nets.modulated_conv2d_fused_act(xp, w, s, noise=nz, up=2, padding=1, resample_filter=f, flip_weight=False, bias=b, act='lrelu', clamp=256.0, out_packed=out)
        for _ in range(2):
            run()
        ts = []
        for _ in range(5):
            flush.zero_()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); c.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(c))
        res_ms[flag] = sorted(ts)[2]
    gf = 2.0 * n * oc * ic * 9 * res * res / 1e9
    print(f'{ic}->{oc} @{res}->{2 * res} n{n} {cg.fp32_precision}: phases + blur {res_ms[True]:.3f} ms ({gf / res_ms[True]:.0f} TF/s alg.)   2x2 4-phase GEMM + blur {res_ms['taps4']:.3f} ms   polyphase {res_ms[False]:.3f} ms ({gf / res_ms[False]:.0f} TF/s alg.)', flush=True)
    del xp, out
