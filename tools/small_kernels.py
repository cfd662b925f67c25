"""One launch of each bandwidth-bound helper kernel at its generator / D-step shape (batch 32), as an ncu target:
    ncu --set full -k regex:'bias_act|fir_|pack_nchw|conv_direct|generic' python tools/small_kernels.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_pkg
load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
up = importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d')
ba = importlib.import_module('pgpp_b200.torch_utils.ops.bias_act')
dev = 'cuda:0'
torch.manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
f = up.setup_filter([1, 3, 3, 1]).to(dev)
cg._init()
with torch.no_grad():
    x = torch.randn(n, 64, 512, 512, device=dev); b = torch.randn(64, device=dev)
    ba.bias_act(x, b, act='lrelu', gain=2 ** 0.5, clamp=256.0)                                  # bias_act_kernel<float>
    xh = x.half()
    ba.bias_act(xh, b.half(), act='lrelu', gain=2 ** 0.5, clamp=256.0)                          # bias_act_kernel<half>
    cg._plugin.pack_activations(x, None, 64, 2)                                                 # pack_nchw_kernel<float, true>
    up.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1])                                            # fir_down2_kernel
    up.upfirdn2d(x, f, padding=[2, 2, 2, 2])                                                    # fir_tile_kernel, 515-wide rows (staged stores)
    up.upfirdn2d(xh, f, padding=[2, 2, 2, 2])                                                   # fir_tile_kernel<half>
    del xh
    x2 = torch.randn(n, 64, 513, 513, device=dev)
    up.upfirdn2d(x2, f, padding=[1, 1, 1, 1], gain=4)                                           # fir_tile_kernel, aligned 512-wide output
    del x2
    x3 = torch.randn(n, 64, 256, 256, device=dev)
    up.upfirdn2d(x3, f, up=2, padding=[2, 1, 2, 1], gain=4)                                     # fir_up2_kernel
    img = torch.randn(n, 3, 256, 256, device=dev)
    up.upfirdn2d(img, f, up=2, padding=[2, 1, 2, 1], gain=4)                                    # fir_up2_kernel, image skip
    par = torch.randint(0, 7, (n, 1, 512, 512), device=dev).float(); w = torch.randn(64, 1, 3, 3, device=dev)
    outp = cg.PackedAct(cg.PackedAct.empty(n, 512, 512, 64, 2, dev), 64)
    cg.direct_conv(par, w, None, wscale=1 / 3, act='relu', out_packed=outp)                     # conv_direct_kernel<3>
    xp = cg.PackedAct(cg._plugin.pack_activations(x, None, 64, 2), 64)
    cg.fir_packed(xp, f, padding=[2, 2, 2, 2])                                                  # fir_tile_packed_kernel
torch.cuda.synchronize()
print('ok')
