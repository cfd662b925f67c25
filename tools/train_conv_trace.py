"""CUDA-event time of every igemm / wgrad launch of ONE training iteration (batch 8), grouped by shape:
    python tools/train_conv_trace.py [top]"""
import collections, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import load_pkg
load_pkg()
ts = importlib.import_module('pgpp_b200.training.training_step')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device('cuda', 0)
torch.manual_seed(0)
G, D, DP = ts.build_networks(dev)
step = ts.TrainingStep(G, D, DP, dev, batch_size=8)
data = bench.train_inputs_to_device(bench.make_train_inputs_u8(8, 200), dev)
for _ in range(2):
    step(data)
torch.cuda.synchronize()
cg.trace = []
step(data)
torch.cuda.synchronize()
by = collections.OrderedDict()
for name, fl, a, b in cg.trace:
    e = by.setdefault(name, [0, 0.0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b); e[2] += fl
cg.trace = None
for kind in ('wgrad', 'igemm'):
    tot = sum(v[1] for k, v in by.items() if k.startswith(kind))
    print(f'{kind}: {tot:.1f} ms in {sum(v[0] for k, v in by.items() if k.startswith(kind))} launches')
for name, (cnt, ms, fl) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{ms:8.3f} ms {cnt:3d}x {fl / ms / 1e9:8.1f} TF/s  {name}')
