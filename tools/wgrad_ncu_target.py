"""one launch each of the weight-gradient kernel on the 64x64 @512^2 n8 layer (tap-pair form, then one tap per accumulator with
PGPP_WGRAD_NO_PAIR=1 in the environment) and on the 128x128 @256^2 n8 layer - the target of the ncu capture in profiles/r02_ncu_wgrad_pair.md"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from conftest import load_pkg  # noqa: E402

load_pkg()
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'
for ca, cb, res in ((64, 64, 512), (128, 128, 256)):
    dy = torch.randn(8, ca, res, res, device=DEV)
    x = torch.randn(8, cb, res, res, device=DEV)
    sp, lp = cg.pack_operand(dy, 'bf16x2'), cg.pack_operand(x, 'bf16x2')
    for flag in ((None, '1') if ca == 64 else (None,)):
        if flag:
            os.environ['PGPP_WGRAD_NO_PAIR'] = flag
        else:
            os.environ.pop('PGPP_WGRAD_NO_PAIR', None)
        custom_ops.refresh_env()
        for _ in range(3):
            cg.weight_gradient(sp, lp, (ca, cb, 3, 3), 1, (1, 1), False, precision='bf16x2', out_dtype=torch.float32)
        torch.cuda.synchronize()
