"""Training-path timing of conv2d_gradfix (BASELINE cfg 5 shapes: discriminator blocks at batch 8 per GPU): forward,
data gradient, weight gradient and the R1 double backward on this package's kernels against the PyTorch library ops
(cuDNN, TF32 off and on).  Prints one line per shape.
    python tools/train_conv_bench.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
dev = 'cuda:0'


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def step(conv, x, w, r1):
    x = x.detach().requires_grad_(True)
    y = conv(x, w, padding=1)
    if r1:      # loss_fullbody.py:264-274: gradient penalty -> second-order graph through the convolution
        gx, = torch.autograd.grad(y.sum(), [x], create_graph=True)
        gx.square().sum().backward()
    else:
        y.square().sum().backward()
    w.grad = None


for (n, c, o, res) in [(8, 64, 64, 512), (8, 64, 128, 256), (8, 128, 256, 128), (8, 256, 512, 64), (8, 512, 512, 32)]:
    x = torch.randn(n, c, res, res, device=dev)
    w = (torch.randn(o, c, 3, 3, device=dev) * 0.05).requires_grad_(True)
    flops = 2.0 * n * res * res * c * o * 9
    row = [f'{c:3d}->{o:3d} k3 {res}x{res} n{n}']
    for r1 in (False, True):
        cg.enabled = True
        ours = timeit(lambda: step(cg.conv2d, x, w, r1))
        cg.enabled = False
        torch.backends.cudnn.allow_tf32 = False
        lib32 = timeit(lambda: step(cg.conv2d, x, w, r1))
        torch.backends.cudnn.allow_tf32 = True
        libtf = timeit(lambda: step(cg.conv2d, x, w, r1))
        cg.enabled = True
        mult = 5 if r1 else 3           # fwd + dgrad + wgrad (+ the two second-order convolutions)
        row.append(f'{"fwd+bwd+R1" if r1 else "fwd+bwd"}: ours (fp32 parity) {ours:.3f} ms = {mult * flops / ours / 1e9:.0f} TF/s alg | '
                   f'library fp32 {lib32:.3f} ms | library TF32 {libtf:.3f} ms')
    print(' ; '.join(row), flush=True)
