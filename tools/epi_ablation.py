"""Ablation timing of one igemm layer (PGPP_IGEMM_DEBUG bits; results are wrong on purpose, only the time matters):
    python tools/epi_ablation.py IC OC RES N PREC [in=packed] [out=packed|nchw]
bits: 1 epilogue without arithmetic / stores, 2 without TMEM loads, 4 no MMAs issued, 8 arithmetic but no stores."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
nets = importlib.import_module('pgpp_b200.training.networks')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
up = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
ic, oc, res, n, prec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
opts = dict(a.split('=') for a in sys.argv[6:])
cg.fp32_precision = prec
parts = cg._PRODUCTS[prec][1]
dev = 'cuda:0'
torch.manual_seed(0)
x = torch.randn(n, ic, res, res, device=dev); w = torch.randn(oc, ic, 3, 3, device=dev); s = torch.randn(n, ic, device=dev)
nz = torch.randn(res, res, device=dev); b = torch.randn(oc, device=dev)
cg._init()
xin = cg.PackedAct(cg._plugin.pack_activations(x, None, ic, parts), ic) if opts.get('in', 'packed') == 'packed' else x
outp = cg.PackedAct(cg.PackedAct.empty(n, res, res, oc, parts, dev), oc) if opts.get('out', 'packed') == 'packed' else None


def run(reps=5):
    cg.trace = []
    with torch.no_grad():
        for _ in range(reps):
            nets.modulated_conv2d_fused_act(xin, w, s, noise=nz, up=1, padding=1, resample_filter=None, flip_weight=True, bias=b,
                                            act='lrelu', gain=2 ** 0.5, clamp=256.0, out_packed=outp)
    torch.cuda.synchronize()
    ts = sorted(t[2].elapsed_time(t[3]) for t in cg.trace)
    return ts[len(ts) // 2], cg.trace[-1][0], cg.trace[-1][1]


AB = {'lean': 'PGPP_IGEMM_NO_LEAN_EPILOGUE', 'tma': 'PGPP_IGEMM_NO_TMA_STORE', 'stack': 'PGPP_IGEMM_NO_STACK', 'slab2': 'PGPP_IGEMM_NO_SLAB2', 'slab9': 'PGPP_IGEMM_SLAB9'}
for env in ([{}, {AB[os.environ['AB']]: '1'}] if os.environ.get('AB') else [{}]):
    for k in AB.values():
        os.environ.pop(k, None)
    os.environ.update(env)
    for dbg in [int(v) for v in os.environ.get('DBG_LIST', '0,8,1,2,4,6,5').split(',')]:
        os.environ['PGPP_IGEMM_DEBUG'] = str(dbg)
        custom_ops.refresh_env()
        ms, name, fl = run()
        print(f'{name} {opts} {env} dbg={dbg}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TF/s', flush=True)
os.environ['PGPP_IGEMM_DEBUG'] = '0'
