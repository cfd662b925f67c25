"""Same-box A/B of a module-level switch on the full generator step (batch 32, CUDA-graph replay, interleaved repetitions):
    python tools/ab_flag.py generator.FUSE_INSTNORM_STATS [reps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
mod_name, attr = sys.argv[1].rsplit('.', 1)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device('cuda', 0)
net = bench.build_generator(dev)
gen = importlib.import_module('pgpp_b200.training.generator')
mod = importlib.import_module('pgpp_b200.training.' + mod_name) if '.' not in mod_name else importlib.import_module(mod_name)
x = bench.to_device_f32(bench.make_generator_inputs_u8(32, 100), dev)
graphs = {}
for val in (True, False):
    setattr(mod, attr, val)
    graphs[val] = gen.GraphedGenerator(net, x)
res = {True: [], False: []}
for r in range(reps):
    for val in (True, False):
        g = graphs[val]
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(8):
            g.replay()
        b.record(); torch.cuda.synchronize()
        res[val].append(a.elapsed_time(b) / 8)
for val in (True, False):
    print(f'{sys.argv[1]} = {val}: ' + ' '.join(f'{t:.2f}' for t in res[val]) + f'  ms/step (median {sorted(res[val])[len(res[val]) // 2]:.2f})')
