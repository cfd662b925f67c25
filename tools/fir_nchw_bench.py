import importlib, os, sys, torch
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
upf = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
f = upf.setup_filter([1, 3, 3, 1], device='cuda')
for dt in (torch.float16, torch.float32):
    x = torch.randn(8, 64, 512, 512, device='cuda').to(dt)
    for name, kw in (('blur pad 2 (before the strided conv)', dict(padding=[2, 2, 2, 2])), ('blur pad 1', dict(padding=[1, 1, 1, 1])), ('down2', dict(down=2, padding=[1, 1, 1, 1])), ('up2', dict(up=2, padding=[2, 1, 2, 1], gain=4))):
        with torch.no_grad():
            y = upf.upfirdn2d(x, f, **kw)
            t = timed(lambda: upf.upfirdn2d(x, f, **kw))
        nbytes = (x.numel() + y.numel()) * x.element_size()
        print(f'{str(dt):14s} 64 ch 512x512 n8 {name:40s} {t:.3f} ms  {nbytes / t / 1e6:.0f} GB/s')
