"""Kernel launches of `__graft_entry__.smoke()` split into this library's kernels and everything else (torch.profiler / CUPTI):
    python tools/smoke_launches.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
from torch.profiler import ProfilerActivity, profile
g.smoke()                                   # first call: library load, weight packing caches
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.smoke()
    torch.cuda.synchronize()
own = other = 0
names = {}
for e in prof.key_averages():
    if e.device_time_total <= 0:
        continue
    k = 'pgpp' if 'pgpp::' in e.key else ('memcpy/memset' if e.key.lower().startswith(('memcpy', 'memset')) else 'other')
    names.setdefault(k, [0, 0.0])
    names[k][0] += e.count; names[k][1] += e.device_time_total / 1e3
for k, (c, ms) in sorted(names.items()):
    print(f'{k:14s} {c:5d} launches  {ms:8.3f} ms')
print('top non-pgpp kernels:')
rows = sorted(((e.count, e.key) for e in prof.key_averages() if e.device_time_total > 0 and 'pgpp::' not in e.key), reverse=True)[:14]
for c, k in rows:
    print(f'  {c:4d}  {k[:120]}')
