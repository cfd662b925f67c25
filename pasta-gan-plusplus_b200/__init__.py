"""B200-native (sm_100a) kernels behind the PASTA-GAN++ / StyleGAN2 synthesis ops.

Package layout mirrors the reference's import paths for the hot path only:
    torch_utils/custom_ops.py            prebuilt C-ABI library loader (replaces the ninja JIT)
    torch_utils/ops/{bias_act,upfirdn2d,conv2d_gradfix,conv2d_resample,fma,grid_sample_gradfix}.py
    training/networks.py                 modulated_conv2d (+ the layer classes that call it)
    csrc/                                CUDA sources of lib/libpgpp_sm100a.so (C ABI: include/pgpp.h)

The directory name is not a Python identifier; load it with `tests/conftest.py:load_pkg()` /
`__graft_entry__.load_pkg()` (module name `pgpp_b200`) or call `install()` to alias the op modules
over an importable reference checkout.
"""
import importlib
import sys

__version__ = '0.1.0'

OP_MODULES = ('bias_act', 'upfirdn2d', 'conv2d_gradfix', 'conv2d_resample', 'fma', 'grid_sample_gradfix')


def install(patch_networks=True, cpu_tensors=None):
    """Drop-in switch: make `torch_utils.ops.<op>` and `torch_utils.custom_ops` resolve to this package
    and replace `training.networks.modulated_conv2d` (if that module is importable / imported).
    Call before the reference model modules are imported.
    cpu_tensors: 'ref' lets the ops take their PyTorch reference path for non-CUDA tensors, which is the reference's own
    dispatch rule (bias_act.py:87, upfirdn2d.py:162) and what its model code relies on when run on the CPU; the default
    ('raise') refuses them.  CUDA tensors always reach the sm_100a kernels."""
    me = sys.modules[__name__]
    if cpu_tensors is not None:
        assert cpu_tensors in ('raise', 'ref')
        importlib.import_module(f'{__name__}.torch_utils.custom_ops').cpu_tensors = cpu_tensors
    ops = importlib.import_module(f'{__name__}.torch_utils.ops')
    for name in OP_MODULES:
        mod = importlib.import_module(f'{__name__}.torch_utils.ops.{name}')
        sys.modules[f'torch_utils.ops.{name}'] = mod
        parent = sys.modules.get('torch_utils.ops')
        if parent is not None:
            setattr(parent, name, mod)
    sys.modules['torch_utils.custom_ops'] = importlib.import_module(f'{__name__}.torch_utils.custom_ops')
    if patch_networks and 'training.networks' in sys.modules:
        nets = importlib.import_module(f'{__name__}.training.networks')
        sys.modules['training.networks'].modulated_conv2d = nets.modulated_conv2d
    return me
