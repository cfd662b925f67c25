"""Batch-1 (and small batch) latency of the full generator: eager launches vs CUDA-graph replay, with a result check."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
gen = importlib.import_module('pgpp_b200.training.generator') if False else None
dev = torch.device('cuda', 0)
G = bench.build_generator(dev)
gen = importlib.import_module('pgpp_b200.training.generator')
for batch in [int(a) for a in (sys.argv[1:] or ['1', '4'])]:
    x = bench.to_device_f32(bench.make_generator_inputs_u8(batch, 100), dev)
    for _ in range(3):
        ref = bench.run_generator(G, x)
    torch.cuda.synchronize()
    def timeit(fn, iters=20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(iters):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    eager = timeit(lambda: bench.run_generator(G, x))
    try:
        gg = gen.GraphedGenerator(G, x)
        out = gg(x)
        torch.cuda.synchronize()
        err = max(float((a - b).abs().max()) for a, b in zip(out, ref))
        graph = timeit(lambda: gg(x))
        print(f'batch {batch}: eager {eager:.2f} ms/step ({batch / eager * 1e3:.1f} img/s), graph {graph:.2f} ms/step ({batch / graph * 1e3:.1f} img/s), max |graph - eager| = {err:.2e}')
    except Exception as e:      # noqa
        print(f'batch {batch}: eager {eager:.2f} ms/step; graph capture failed: {type(e).__name__}: {str(e)[:300]}')
