"""Two full-generator passes at batch 32 (first pass warms caches / packs weights): target for the ncu launch list."""
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
print('done')
