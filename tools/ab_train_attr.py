"""Same-box A/B of a module-level switch on training phases (BASELINE configs[4], batch 8, 1 GPU; interleaved repetitions, CUDA events):
    python tools/ab_train_attr.py generator.MERGE_GAMMA_BETA [reps] [phase ...]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import load_pkg

load_pkg()
ts = importlib.import_module('pgpp_b200.training.training_step')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
mod_name, attr = sys.argv[1].rsplit('.', 1)
mod = importlib.import_module(('pgpp_b200.training.' if mod_name in ('generator', 'synthesis', 'networks', 'discriminator') else 'pgpp_b200.torch_utils.ops.') + mod_name)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
phases = sys.argv[3:] or None
dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
G, D, DP = ts.build_networks(dev)
step = ts.TrainingStep(G, D, DP, dev, batch_size=8)
data = bench.train_inputs_to_device(bench.make_train_inputs_u8(8, 200), dev)
res = {True: [], False: []}
launches = {}
for r in range(reps + 1):
    for v in (True, False):
        setattr(mod, attr, v)
        step(data, phases=phases)
        torch.cuda.synchronize()
        l0 = custom_ops.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(data, phases=phases)
        b.record()
        torch.cuda.synchronize()
        launches[v] = custom_ops.launch_count() - l0
        if r:
            res[v].append(a.elapsed_time(b))
for v in (True, False):
    xs = sorted(res[v])
    print(f'{sys.argv[1]} = {v}: median {xs[len(xs) // 2]:.1f} ms  (min {xs[0]:.1f}, max {xs[-1]:.1f}; {launches[v]} launches through the C ABI)  phases {phases or "all"}')
