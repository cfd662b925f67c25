"""Kernel table of ONE training iteration (BASELINE configs[4], batch 8, 1 GPU) from torch.profiler (CUPTI kernel records):
device time per kernel name, split into this repo's kernels (pgpp::), library GEMM / reductions and ATen element-wise kernels.
    python tools/train_trace.py [batch] [phase ...]        # default: all phases"""
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
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
phases = sys.argv[2:] or None
dev = torch.device('cuda', 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
G, D, DP = ts.build_networks(dev)
step = ts.TrainingStep(G, D, DP, dev, batch_size=batch)
data = bench.train_inputs_to_device(bench.make_train_inputs_u8(batch, 200), dev)
for _ in range(2):
    step(data, phases=phases)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
l0 = custom_ops.launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    a.record()
    step(data, phases=phases)
    b.record()
    torch.cuda.synchronize()
native_launches = custom_ops.launch_count() - l0
rows = {}
for e in prof.events():
    if e.device_type is not None and str(e.device_type).endswith('CUDA') and e.device_time > 0:
        r = rows.setdefault(e.name, [0.0, 0])
        r[0] += e.device_time / 1e3
        r[1] += 1


def kind(name):
    if name.startswith('pgpp::') or 'pgpp::' in name:
        return 'pgpp'
    if 'nccl' in name.lower():
        return 'nccl'
    if any(k in name for k in ('gemm', 'cutlass', 'cublas', 'sgemm', 'nvjet', 'gemv')):
        return 'library gemm'
    if 'Memcpy' in name or 'Memset' in name:
        return 'memcpy/memset'
    return 'aten / other'


tot = sum(r[0] for r in rows.values())
by = {}
for n, (ms, c) in rows.items():
    k = by.setdefault(kind(n), [0.0, 0]); k[0] += ms; k[1] += c
print(f'one training iteration, batch {batch}, phases {phases or "all"}: {a.elapsed_time(b):.1f} ms wall (under the profiler), '
      f'{tot:.1f} ms of kernel time in {sum(r[1] for r in rows.values())} launches, {native_launches} of them through the C ABI')
for k, (ms, c) in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print(f'  {k:16s} {ms:9.2f} ms  {100 * ms / tot:5.1f} %  {c:6d} launches')
print()
for n, (ms, c) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:60]:
    print(f'{ms:9.3f} ms {100 * ms / tot:5.1f} % {c:6d}  [{kind(n):13s}] {n[:150]}')
