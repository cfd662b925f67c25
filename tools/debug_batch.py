"""Find the first module whose output for sample 0 differs between a batch-1 and a batch-2 run of the generator."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
from oracle import ref_generator
load_pkg()
gen = importlib.import_module('pgpp_b200.training.generator')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
DEV = 'cuda:0'
G = gen.build_generator().eval()
ref_generator.name_seeded_init(list(G.named_parameters()) + list(G.named_buffers()))
G = G.to(DEV).requires_grad_(False)
inp = {k: v.to(DEV) for k, v in ref_generator.synthetic_inputs(1, seed=0).items()}
inp2 = {k: torch.cat([v, v.flip(-1)]) for k, v in inp.items()}
rec = {}
def hook(name, store):
    def fn(mod, args, out):
        outs = out if isinstance(out, (tuple, list)) else (out,)
        vals = []
        for o in outs:
            if isinstance(o, cg.PackedAct):
                vals.append(o.to_nchw()[:1].clone())
            elif torch.is_tensor(o):
                vals.append(o[:1].detach().clone())
            elif isinstance(o, (tuple, list)):
                vals += [t[:1].detach().clone() for t in o if torch.is_tensor(t)]
        store.setdefault(name, []).append(vals)
    return fn
def run(x, store):
    hs = [m.register_forward_hook(hook(n, store)) for n, m in G.named_modules() if n and n.count('.') <= 2]
    with torch.no_grad():
        out = G(torch.zeros(x['c'].shape[0], 0, device=DEV), x['c'], x['retain'], x['pose'], x['denorm_upper'], x['denorm_lower'],
                x['denorm_upper_mask'], x['denorm_lower_mask'], gt_parsing=x['gt_parsing'], noise_mode='const')
    for h in hs: h.remove()
    return out
r1, r2 = {}, {}
run(inp, r1); run(inp2, r2)
bad = 0
for name in r1:
    for call, (a, b) in enumerate(zip(r1[name], r2[name])):
        for i, (x, y) in enumerate(zip(a, b)):
            if x.shape != y.shape: continue
            e = float((x.float() - y.float()).norm() / x.float().norm().clamp_min(1e-20))
            if e > 1e-5:
                d = (x.float() - y.float()).abs()
                rows = d.amax(dim=(0, 1, 3)); nz = (rows > 1e-4 * float(x.abs().max())).nonzero().flatten()
                print(f'{name} call{call} out{i} shape={tuple(x.shape)} rel={e:.2e} bad rows {nz[:3].tolist()}..{nz[-3:].tolist()} ({len(nz)})')
                bad += 1
                if bad > 12: sys.exit(0)
print('done, mismatches:', bad)
