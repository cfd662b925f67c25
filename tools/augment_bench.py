"""grid_sample and AugmentPipe timing at ADA sizes (discriminator input, 512 x 512 RGB):
    python tools/augment_bench.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
load_pkg()
gs = importlib.import_module('pgpp_b200.torch_utils.ops.grid_sample_gradfix')
augment = importlib.import_module('pgpp_b200.training.augment')
dev = 'cuda:0'
HBM = 6549.1


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in (8, 32):
    # the pipeline samples the 2x supersampled, padded image: (512 + 2 * 3 + margins) * 2 ~ 1100 px
    x = torch.randn(n, 3, 1100, 1100, device=dev)
    theta = torch.tensor([[0.9, 0.2, 0.05], [-0.2, 0.9, -0.03]], device=dev).repeat(n, 1, 1)
    grid = torch.nn.functional.affine_grid(theta, [n, 3, 1036, 1036], align_corners=False)
    dy = torch.randn(n, 3, 1036, 1036, device=dev)
    nbytes = (x.numel() + grid.numel() + dy.numel()) * 4
    ours = timeit(lambda: gs.grid_sample(x, grid))
    lib = timeit(lambda: torch.nn.functional.grid_sample(x, grid, mode='bilinear', padding_mode='zeros', align_corners=False))
    print(f'grid_sample fwd n={n}: ours {ours:.3f} ms ({nbytes / ours / 1e6:.0f} GB/s, {nbytes / ours / 1e6 / HBM * 100:.0f}% of HBM copy) | library {lib:.3f} ms', flush=True)
    plugin = gs._plugin
    ours_b = timeit(lambda: plugin.backward(dy, x, grid))
    op = torch.ops.aten.grid_sampler_2d_backward
    lib_b = timeit(lambda: op(dy, x, grid, 0, 0, False, [True, True]))
    print(f'grid_sample bwd n={n}: ours {ours_b:.3f} ms | library {lib_b:.3f} ms', flush=True)
bgc = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1)
pipe = augment.AugmentPipe(**bgc).to(dev)
pipe.p.copy_(torch.as_tensor(0.6))
for n in (8, 32):
    img = torch.randn(n, 3, 512, 512, device=dev).clamp(-1, 1)
    ms = timeit(lambda: pipe(img), reps=5)
    print(f'AugmentPipe bgc p=0.6, {n} x 3 x 512 x 512: {ms:.3f} ms = {n / ms * 1e3:.0f} images/s', flush=True)
