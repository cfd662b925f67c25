"""Bring-up probe for csrc/conv_igemm.cu: one case per process (a trapped kernel kills the CUDA context).
    python tools/igemm_probe.py <case index | all>
Prints relative-L2 / max-abs error against a float64 CPU emulation fed with the SAME bf16-rounded operands,
so anything above ~1e-5 is a kernel bug, not quantisation."""
import importlib
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))

CASES = [
    # name, N, C, H, W, O, k, pad, stride, up, products, extras
    ('gemm_c64_o64',      1, 64, 16, 16, 64, 1, 0, 1, 1, 1, {}),
    ('gemm_c128_o128',    1, 128, 16, 16, 128, 1, 0, 1, 1, 1, {}),
    ('conv3_c64_o64',     1, 64, 16, 16, 64, 3, 1, 1, 1, 1, {}),
    ('conv3_c16_o16',     2, 16, 16, 16, 16, 3, 1, 1, 1, 1, {}),
    ('conv3_c32_o32',     2, 32, 16, 16, 32, 3, 1, 1, 1, 1, {}),
    ('conv3_c48_o24',     2, 48, 20, 12, 24, 3, 1, 1, 1, 1, {}),
    ('many_tiles_o256',   4, 64, 64, 64, 256, 3, 1, 1, 1, 1, {}),
    ('o512_two_coltiles', 8, 128, 32, 32, 512, 3, 1, 1, 1, 1, {}),
    ('split3',            2, 64, 16, 16, 64, 3, 1, 1, 1, 3, {}),
    ('split6',            2, 64, 16, 16, 64, 3, 1, 1, 1, 6, {}),
    ('up2_c64_o64',       2, 64, 16, 16, 64, 3, 1, 1, 2, 1, {}),
    ('up2_c32_o16',       2, 32, 8, 8, 16, 3, 1, 1, 2, 6, {}),
    ('stride2',           2, 64, 32, 32, 64, 3, 1, 2, 1, 1, {}),
    ('tiny_8x8_tn2',      4, 64, 8, 8, 64, 3, 1, 1, 1, 1, {}),
    ('tiny_4x4_tn8',      16, 64, 4, 4, 64, 3, 1, 1, 1, 1, {}),
    ('torgb_o3',          2, 64, 32, 32, 3, 1, 0, 1, 1, 1, {}),
    ('k7_c3',             1, 3, 32, 32, 64, 7, 3, 1, 1, 6, {}),
    ('epilogue_full',     2, 64, 16, 16, 64, 3, 1, 1, 1, 6, {'epi': True}),
    ('bf16_nhwc_out',     2, 64, 16, 16, 64, 3, 1, 1, 1, 1, {'nhwc_bf16': True}),
    ('wide_128px_rows',   1, 64, 8, 256, 64, 3, 1, 1, 1, 1, {}),
    ('resident_64_bf16x2', 4, 64, 128, 128, 64, 3, 1, 1, 1, 3, {}),
    ('resident_64_bf16',  4, 64, 128, 128, 64, 3, 1, 1, 1, 1, {}),
    ('ring_128_bf16x2',   2, 128, 128, 128, 128, 3, 1, 1, 1, 3, {}),
    ('ragged_reuse',      3, 48, 37, 53, 40, 3, 1, 1, 1, 6, {}),
    ('k5_reuse',          2, 32, 24, 40, 32, 5, 2, 1, 1, 3, {}),
    ('k7_c3_big',         2, 3, 96, 64, 64, 7, 3, 1, 1, 3, {}),
    ('up2_many_tiles',    4, 128, 64, 64, 64, 3, 1, 1, 2, 3, {}),
    ('torgb_resident',    4, 64, 128, 128, 3, 1, 0, 1, 1, 3, {}),
]


def run_case(idx):
    from conftest import load_pkg
    from helpers import emulate_igemm, rel_l2, max_abs
    load_pkg()
    cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
    up_mod = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
    name, n, c, h, w, o, k, pad, stride, up, products, extra = CASES[idx]
    parts = {1: 1, 3: 2, 6: 3}[products]
    prec = {1: 'bf16', 3: 'bf16x2', 6: 'bf16x3'}[products]
    g = torch.Generator().manual_seed(100 + idx)
    x = torch.randn(n, c, h, w, generator=g)
    wt = torch.randn(o, c, k, k, generator=g) / (c * k * k) ** 0.5
    dev = 'cuda:0'
    if up == 2:
        f = up_mod.setup_filter([1, 3, 3, 1])
        pw = cg.packed_up2(wt.to(dev), f.to(dev), False, False, parts)
    else:
        pw = cg.packed_plain(wt.to(dev), True, parts, pad, pad)
    # reference operands: exactly what the kernel sees
    xr = cg._split_bf16(x, parts).double().sum(0)
    want = emulate_igemm(xr, pw, stride=stride, parts=parts)
    kw = {}
    xin = x.to(dev)
    if extra.get('epi'):
        d = torch.rand(n, o, generator=g) + 0.5
        nz = torch.randn(h * up, w * up, generator=g)
        b = torch.randn(o, generator=g)
        kw = dict(dcoef=d.to(dev), noise=nz.to(dev), bias=b.to(dev), act='lrelu', alpha=0.2, gain=2 ** 0.5, clamp=1.5)
        v = want * d.double().reshape(n, o, 1, 1) + nz.double() + b.double().reshape(1, o, 1, 1)
        want = (torch.where(v > 0, v, v * 0.2) * 2 ** 0.5).clamp(-1.5, 1.5)
    if extra.get('nhwc_bf16'):
        kw = dict(out_dtype=torch.bfloat16, memory_format=torch.channels_last)
    got = cg.igemm_conv(xin, pw, stride=stride, precision=prec, **kw)
    torch.cuda.synchronize()
    r, m = rel_l2(got, want), max_abs(got, want)
    tol = 1e-2 if extra.get('nhwc_bf16') else 2e-5
    status = 'OK ' if r < tol else 'BAD'
    print(f'{status} {name:20s} rel={r:.3e} maxabs={m:.3e} shape={tuple(got.shape)}', flush=True)
    if r >= tol:
        e = (got.double().cpu() - want).abs()
        print('   err by channel (first 16):', [f'{v:.2e}' for v in e.amax(dim=(0, 2, 3))[:16].tolist()])
        print('   err by row     (first 16):', [f'{v:.2e}' for v in e.amax(dim=(0, 1, 3))[:16].tolist()])
        print('   err by col     (first 16):', [f'{v:.2e}' for v in e.amax(dim=(0, 1, 2))[:16].tolist()])
        print('   got[0,0,0,:8] ', got[0, 0, 0, :8].float().cpu().tolist())
        print('   want[0,0,0,:8]', want[0, 0, 0, :8].tolist())
    return 0 if r < tol else 1


if __name__ == '__main__':
    arg = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if arg == 'all':
        bad = 0
        for i in range(len(CASES)):
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), str(i)], timeout=120, capture_output=True, text=True)
                out = (p.stdout + p.stderr).strip().splitlines()
                keep = [l for l in out if l.startswith(('OK', 'BAD', '   '))] or out[-6:]
                print('\n'.join(keep), flush=True)
                bad += p.returncode != 0
            except subprocess.TimeoutExpired:
                print(f'TIMEOUT case {i} {CASES[i][0]}', flush=True)
                bad += 1
        print(f'probe: {len(CASES) - bad}/{len(CASES)} cases ok')
        sys.exit(1 if bad else 0)
    sys.exit(run_case(int(arg)))
