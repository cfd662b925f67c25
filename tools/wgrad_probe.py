"""GPU probe of pgpp_conv2d_wgrad: each case in its own subprocess (a broken pipeline traps instead of hanging the
box, and a trap poisons only that process), checked against float64 autograd on the CPU, timed against the
library's weight gradient.

    python tools/wgrad_probe.py            # all cases
    python tools/wgrad_probe.py 3          # one case, in-process
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (n, o, i, h, w, k, stride, pad, transpose, precision)
CASES = [
    (2, 64, 64, 16, 64, 3, 1, 1, False, 'bf16x2'),      # bw = 64, one sample row per K block
    (2, 64, 64, 16, 16, 3, 1, 1, False, 'bf16x2'),      # bw = 16 (inner TMA box 32 bytes)
    (2, 32, 48, 8, 8, 3, 1, 1, False, 'bf16x2'),        # bw = 8, ragged channels
    (3, 16, 16, 4, 4, 3, 1, 1, False, 'bf16x3'),        # 4x4 images (row pitch padded to 8)
    (2, 128, 64, 32, 32, 1, 1, 0, False, 'bf16x2'),     # 1x1
    (2, 64, 64, 32, 32, 3, 2, 1, False, 'bf16x2'),      # stride 2 (column-parity planes)
    (2, 64, 32, 33, 33, 3, 2, 0, False, 'bf16x2'),      # stride 2 on the blurred odd-size image (conv2d_resample down=2)
    (2, 32, 64, 16, 16, 3, 2, 0, True, 'bf16x2'),       # conv_transpose2d stride 2 (up=2 layer)
    (2, 64, 64, 16, 16, 3, 1, 1, True, 'bf16x2'),       # conv_transpose2d stride 1 (data gradient of a conv)
    (1, 3, 64, 64, 64, 1, 1, 0, False, 'bf16x2'),       # ToRGB
    (2, 64, 3, 32, 32, 7, 1, 3, False, 'bf16x2'),       # 7x7 RGB stem
    (4, 256, 256, 32, 32, 3, 1, 1, False, 'bf16'),      # several a/b blocks, bf16 mode
    (4, 512, 512, 16, 16, 3, 1, 1, False, 'bf16x2'),
    (8, 128, 128, 128, 128, 3, 1, 1, False, 'bf16x2'),  # timing
    (8, 512, 512, 32, 32, 3, 1, 1, False, 'bf16x2'),    # timing
    (8, 64, 64, 256, 256, 3, 1, 1, False, 'bf16x2'),    # timing
    (8, 512, 512, 32, 32, 3, 1, 1, False, 'bf16'),      # timing
]
TOL = {'bf16x3': 4e-5, 'bf16x2': 1e-4, 'bf16': 1.5e-2}


def run_case(idx):
    import torch
    import torch.nn.functional as F
    from __graft_entry__ import load_pkg
    pkg = load_pkg()
    from pgpp_b200.torch_utils.ops import conv2d_gradfix
    n, o, i, h, w, k, s, p, tr, prec = CASES[idx]
    g = torch.Generator().manual_seed(idx)
    x = torch.randn(n, i, h, w, generator=g)
    wshape = (i, o, k, k) if tr else (o, i, k, k)
    wt = torch.randn(*wshape, generator=g, dtype=torch.float64, requires_grad=True)
    if tr:
        y = F.conv_transpose2d(x.double(), wt, stride=s, padding=p)
    else:
        y = F.conv2d(x.double(), wt, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=g)
    ref = torch.autograd.grad(y, wt, dy.double())[0]
    xd, dyd = x.cuda(), dy.cuda()
    got = conv2d_gradfix.weight_gradient(dyd, xd, wshape, s, (p, p), tr, precision=prec)
    torch.cuda.synchronize()
    err = ((got.cpu().double() - ref).norm() / ref.norm()).item()
    ok = err <= TOL[prec]
    msg = f'case {idx:2d} {CASES[idx]}: rel-L2 {err:.2e} {"ok" if ok else "FAIL"}'
    if n >= 8:
        def timeit(fn, reps=10):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        conv2d_gradfix.trace = []
        conv2d_gradfix.weight_gradient(dyd, xd, wshape, s, (p, p), tr, precision=prec)
        torch.cuda.synchronize()
        kern_ms = conv2d_gradfix.trace[0][2].elapsed_time(conv2d_gradfix.trace[0][3])
        conv2d_gradfix.trace = None
        mine = timeit(lambda: conv2d_gradfix.weight_gradient(dyd, xd, wshape, s, (p, p), tr, precision=prec))
        wd = torch.empty(wshape, device='cuda')
        lib = timeit(lambda: torch.ops.aten.convolution_backward(dyd, xd, wd, None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [False, True, False]))
        xb, dyb, wb = xd.bfloat16(), dyd.bfloat16(), wd.bfloat16()
        libb = timeit(lambda: torch.ops.aten.convolution_backward(dyb, xb, wb, None, [s, s], [p, p], [1, 1], tr, [0, 0], 1, [False, True, False]))
        flops = 2.0 * n * y.shape[2] * y.shape[3] * o * i * k * k
        msg += (f' | kernel {kern_ms:.3f} ms ({flops / kern_ms / 1e9:.0f} TF/s alg), with split {mine:.3f} ms;'
                f' library fp32 {lib:.3f} ms, library bf16 {libb:.3f} ms')
    print(msg, flush=True)
    return ok


if __name__ == '__main__':
    if len(sys.argv) > 1:
        sys.exit(0 if run_case(int(sys.argv[1])) else 1)
    bad = 0
    for idx in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(idx)], timeout=180, capture_output=True, text=True)
            out = (r.stdout.strip().splitlines() or ['<no output>'])[-1]
            if r.returncode != 0:
                bad += 1
                out += ' || rc=%d %s' % (r.returncode, r.stderr.strip()[-400:].replace('\n', ' / '))
            print(out, flush=True)
        except subprocess.TimeoutExpired:
            bad += 1
            print(f'case {idx}: TIMEOUT', flush=True)
    print(f'{len(CASES) - bad}/{len(CASES)} cases ok')
    sys.exit(1 if bad else 0)
