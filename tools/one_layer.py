"""Run one modulated conv layer a few times (target for ncu / quick timing):
    python tools/one_layer.py IC OC RES UP N PREC [reps] [in=tensor|packed] [out=nchw|packed]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
nets = importlib.import_module('pgpp_b200.training.networks')
up = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
ic, oc, res, upf, n, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
opts = dict(a.split('=') for a in sys.argv[8:])
cg.fp32_precision = prec
parts = cg._PRODUCTS[prec][1]
dev = 'cuda:0'
torch.manual_seed(0)
x = torch.randn(n, ic, res, res, device=dev); w = torch.randn(oc, ic, 3, 3, device=dev); s = torch.randn(n, ic, device=dev)
nz = torch.randn(res * upf, res * upf, device=dev); b = torch.randn(oc, device=dev); f = up.setup_filter([1, 3, 3, 1]).to(dev)
xin = x
if opts.get('in') == 'packed':
    cg._init()
    xin = cg.PackedAct(cg._plugin.pack_activations(x, None, ic, parts), ic)
outp = None
if opts.get('out') == 'packed':
    outp = cg.PackedAct(cg.PackedAct.empty(n, res * upf, res * upf, oc, parts, dev), oc)
cg.trace = []
with torch.no_grad():
    for _ in range(reps):
        y = nets.modulated_conv2d_fused_act(xin, w, s, noise=nz, up=upf, padding=1, resample_filter=f, flip_weight=(upf == 1), bias=b,
                                            act='lrelu', gain=2 ** 0.5, clamp=256.0, out_packed=outp)
torch.cuda.synchronize()
t = cg.trace[-1]
ms = t[2].elapsed_time(t[3])
print(f'{t[0]} {opts}: {ms:.3f} ms, {t[1] / ms / 1e9:.1f} TFLOP/s')
