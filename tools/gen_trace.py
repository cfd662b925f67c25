"""Per-launch CUDA-event times of every igemm launch of one full-generator step, grouped by shape:
    python tools/gen_trace.py [batch] [precision]"""
import collections, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16x2'
dev = torch.device('cuda', 0)
G = bench.build_generator(dev)
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
cg.fp32_precision = prec
x = bench.to_device_f32(bench.make_generator_inputs_u8(batch, 100), dev)
for _ in range(2):
    bench.run_generator(G, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cg.trace = []
e0.record(); bench.run_generator(G, x); e1.record()
torch.cuda.synchronize()
by = collections.OrderedDict()
for name, fl, a, b in cg.trace:
    e = by.setdefault(name, [0, 0.0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b); e[2] += fl
tot = sum(v[1] for v in by.values())
for name, (cnt, ms, fl) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f'{ms:8.3f} ms {cnt:3d}x {fl / ms / 1e9:8.1f} TF/s  {name}')
print(f'igemm total {tot:.2f} ms ({len(cg.trace)} launches) of step {e0.elapsed_time(e1):.2f} ms')
