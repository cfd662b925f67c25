"""pgpp_fir_packed against the HBM roofline (bytes = read + write of the operand-format tensor):
    python tools/fir_packed_bench.py"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
up = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
dev = 'cuda:0'
peak = 6541.5
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
f = up.setup_filter([1, 3, 3, 1]).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, c, h, down, pad in [(32, 64, 512, 1, (2, 2, 2, 2)), (32, 128, 256, 1, (2, 2, 2, 2)), (32, 256, 128, 1, (2, 2, 2, 2)), (32, 64, 256, 1, (2, 2, 2, 2)),
                           (32, 64, 512, 2, (1, 1, 1, 1)), (32, 64, 512, 1, None)]:
    x = cg.PackedAct(cg.PackedAct.empty(n, h, h, c, 2, dev), c)
    x.data.normal_()
    ff = None if pad is None else f
    pad = pad or (0, 0, 0, 0)
    out = cg.fir_packed(x, ff, down=down, padding=pad)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cg.fir_packed(x, ff, down=down, padding=pad, out=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = x.data.numel() * 2 + out.data.numel() * 2
    print(f'fir_packed n{n} c{c} {h}x{h} down{down} {"blur 4x4" if ff is not None else "slice copy"}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  {nbytes / ms / 1e6 / peak:.2f} of measured HBM copy')
