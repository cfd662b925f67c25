"""torch.profiler kernel table of one generator step (batch from argv, default 32): where the non-igemm time goes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda', 0)
G = bench.build_generator(dev)
x = bench.to_device_f32(bench.make_generator_inputs_u8(batch, 100), dev)
for _ in range(2):
    bench.run_generator(G, x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    bench.run_generator(G, x)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows if not r[0].startswith(('aten::', 'cudaLaunch', 'cuda')))
print('sum of kernel rows (ms):', round(tot, 2))
for k, ms, c in rows[:45]:
    print(f'{ms:9.3f} ms {c:5d}  {k[:110]}')
