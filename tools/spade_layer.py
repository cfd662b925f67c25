"""Run one fused SPADE normalisation (conv_mlp -> gamma|beta GEMM with the SPADE epilogue) a few times (target for ncu / quick timing):
    python tools/spade_layer.py C RES N PREC [reps]          # C = normalised channels (64 @512, 128 @256 in the generator)"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
gen = importlib.import_module('pgpp_b200.training.generator')
c, res, n, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cg.fp32_precision = prec
dev = 'cuda:0'
torch.manual_seed(0)
blk = gen.Spade_Norm_Block(1, c).to(dev).eval()
x = torch.randn(n, c, res, res, device=dev)
parsing = torch.randint(0, 7, (n, 1, res, res), device=dev).float()
var, mean = torch.var_mean(x, dim=(2, 3), unbiased=False)
rstd = (var + 1e-5).rsqrt()
cg.trace = []
with torch.no_grad():
    for _ in range(reps):
        y = blk.fused_packed(x, mean, rstd, gen.RawFeat(parsing), 2 ** 0.5)
torch.cuda.synchronize()
for t in cg.trace[-1:]:
    ms = t[2].elapsed_time(t[3])
    print(f'{t[0]}: {ms:.3f} ms, {t[1] / ms / 1e9:.1f} TFLOP/s')
