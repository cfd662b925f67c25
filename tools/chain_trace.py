"""Print every igemm launch of one synthesis-chain step with its CUDA-event time: python tools/chain_trace.py [batch] [precision]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else 'bf16x2'
dev = torch.device('cuda', 0)
net = bench.build_chain(dev)
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
cg.fp32_precision = prec
cat = bench.make_cat_feats(net, batch, dev)
ws, pose = bench.make_inputs(net, batch, 100)
ws, pose = ws.to(dev), pose.to(dev)
with torch.no_grad():
    for _ in range(2):
        net(ws, pose, cat, noise_mode='const')
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cg.trace = []
    e0.record()
    net(ws, pose, cat, noise_mode='const')
    e1.record()
    torch.cuda.synchronize()
tot = 0
for t in cg.trace:
    ms = t[2].elapsed_time(t[3]); tot += ms
    print(f'{ms:8.3f} ms {t[1] / ms / 1e9:8.1f} TF/s  {t[0]}')
print(f'igemm total {tot:.2f} ms of step {e0.elapsed_time(e1):.2f} ms')
