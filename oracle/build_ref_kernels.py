"""TEST / BASELINE INFRASTRUCTURE, not product code.  Compiles the reference's OWN CUDA plugins for sm_100a, from the sources where
they lie under /root/reference (nothing is copied into the repo), into oracle/_ref/ (git-ignored, travels to the GPU box):

    /root/reference/torch_utils/ops/bias_act.cpp + bias_act.cu      -> oracle/_ref/ref_bias_act_plugin/ref_bias_act_plugin.so
    /root/reference/torch_utils/ops/upfirdn2d.cpp + upfirdn2d.cu    -> oracle/_ref/ref_upfirdn2d_plugin/ref_upfirdn2d_plugin.so

(the reference's own loader, custom_ops.get_plugin, cannot import what it builds on torch 2.x, so the
modules are loaded by path.)  tools/ref_kernels_baseline.py times them on the B200 next to this repo's kernels - the GPU baseline of
the reference's kernels recompiled, which is what the new kernels have to beat.

    python oracle/build_ref_kernels.py
"""
import os
import sys

REF_OPS = '/root/reference/torch_utils/ops'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
PLUGINS = {'ref_bias_act_plugin': ['bias_act.cpp', 'bias_act.cu'], 'ref_upfirdn2d_plugin': ['upfirdn2d.cpp', 'upfirdn2d.cu']}


def so_path(name):
    return os.path.join(OUT, name, name + '.so')


def build(verbose=False):
    if not os.path.isdir(REF_OPS):
        print(f'{REF_OPS} not present: keeping the prebuilt files under {OUT}')
        return False
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    import torch.utils.cpp_extension as ext
    for name, srcs in PLUGINS.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        newest = max(os.path.getmtime(os.path.join(REF_OPS, s)) for s in srcs)
        if os.path.isfile(so_path(name)) and os.path.getmtime(so_path(name)) >= newest:
            continue
        ext.load(name=name, sources=[os.path.join(REF_OPS, s) for s in srcs], build_directory=bdir, verbose=verbose,
                 extra_cuda_cflags=["--use_fast_math", "-lineinfo"])      # the flags the reference passes (bias_act.py:48, upfirdn2d.py:32)
        assert os.path.isfile(so_path(name)), so_path(name)
        print('built', so_path(name))
    return True


def load(name):
    """import a prebuilt plugin by path (GPU box: /root/reference does not exist there)"""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, so_path(name))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    build(verbose='-v' in sys.argv)
