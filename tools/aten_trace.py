"""Library (aten) kernels left in one generator pass, grouped by op and input shape (torch.profiler):
    python tools/aten_trace.py [batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda', 0)
G = bench.build_generator(dev)
x = bench.to_device_f32(bench.make_generator_inputs_u8(batch, 100), dev)
for _ in range(2):
    bench.run_generator(G, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    bench.run_generator(G, x)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, 'self_device_time_total', None)
    if t is None:
        t = e.self_cuda_time_total
    if t > 50:
        rows.append((t / 1e3, e.count, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f'total device time of aten ops (self): {tot:.2f} ms')
for ms, cnt, key, shp in rows[:45]:
    print(f'{ms:8.3f} ms {cnt:4d}x {key:38s} {shp}')
