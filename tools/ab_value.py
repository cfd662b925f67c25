"""Same-box A/B of a module-level VALUE on the full generator step (batch 32, CUDA-graph replay, interleaved repetitions):
    python tools/ab_value.py networks.UP2_TAPS4_MIN_IO 8192 32768 [reps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
mod_name, attr = sys.argv[1].rsplit('.', 1)
vals = [int(sys.argv[2]), int(sys.argv[3])]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device('cuda', 0)
net = bench.build_generator(dev)
gen = importlib.import_module('pgpp_b200.training.generator')
mod = importlib.import_module('pgpp_b200.training.' + mod_name)
x = bench.to_device_f32(bench.make_generator_inputs_u8(32, 100), dev)
graphs = {}
for v in vals:
    setattr(mod, attr, v)
    graphs[v] = gen.GraphedGenerator(net, x)
res = {v: [] for v in vals}
for r in range(reps):
    for v in vals:
        g = graphs[v]
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(8):
            g.replay()
        b.record(); torch.cuda.synchronize()
        res[v].append(a.elapsed_time(b) / 8)
for v in vals:
    print(f'{sys.argv[1]} = {v}: ' + ' '.join(f'{t:.2f}' for t in res[v]) + f'  ms/step (median {sorted(res[v])[len(res[v]) // 2]:.2f})')
