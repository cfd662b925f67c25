"""Mint golden fixtures from the REAL reference (TEST INFRASTRUCTURE ONLY; runs only in the build
container, where /root/reference is mounted).

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

It imports the unmodified reference ops from /root/reference (`torch_utils.ops.*`, `impl='ref'`
path on CPU tensors) and `training.networks.modulated_conv2d` through the import shims of
SURVEY.md Appendix E (stub matplotlib, cwd = reference root, spoofed torch.version.cuda), runs them
on small seeded inputs and stores inputs + outputs.  Nothing under tests/, bench.py or smoke()
reads /root/reference at run time; they read the .npz files written here.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


@contextlib.contextmanager
def reference_imports():
    """Make `import torch_utils...` / `import training.networks` resolve to the reference."""
    saved_path = list(sys.path)
    saved_cwd = os.getcwd()
    saved_cuda = torch.version.cuda
    for name in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    os.chdir(REF)
    torch.version.cuda = '11.0'
    try:
        yield
    finally:
        torch.version.cuda = saved_cuda
        os.chdir(saved_cwd)
        sys.path[:] = saved_path


def rnd(gen, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen) * scale


def golden_bias_act(ref_bias_act):
    g = torch.Generator().manual_seed(1234)
    out = {}
    cases = []
    for act in ref_bias_act.activation_funcs.keys():
        cases.append((act, None, None, None))
        cases.append((act, 0.3, 1.7, 0.9))
    cases += [('lrelu', None, float(np.sqrt(2)), 256.0), ('linear', None, 1.0, 256.0), ('linear', None, float(np.sqrt(0.5)), None),
              ('relu', None, 1.0, None)]
    x = rnd(g, 2, 5, 6, 7, scale=2.0)
    b = rnd(g, 5)
    out['x'] = x.numpy()
    out['b'] = b.numpy()
    for i, (act, alpha, gain, clamp) in enumerate(cases):
        xi = x.clone().requires_grad_(True)
        bi = b.clone().requires_grad_(True)
        y = ref_bias_act.bias_act(xi, bi, dim=1, act=act, alpha=alpha, gain=gain, clamp=clamp, impl='ref')
        dy = rnd(torch.Generator().manual_seed(77 + i), *y.shape)
        dx, db = torch.autograd.grad(y, [xi, bi], dy, create_graph=True)
        # second order: derivative of <dx, v> w.r.t. dy-path input x (only meaningful for smooth acts)
        v = rnd(torch.Generator().manual_seed(99 + i), *y.shape)
        if dx.requires_grad:
            ddx, = torch.autograd.grad(dx, [xi], v, allow_unused=True)
        else:
            ddx = None
        tag = f'case{i}'
        out[f'{tag}_meta'] = np.array([act, str(alpha), str(gain), str(clamp)])
        out[f'{tag}_y'] = y.detach().numpy()
        out[f'{tag}_dy'] = dy.numpy()
        out[f'{tag}_dx'] = dx.detach().numpy()
        out[f'{tag}_db'] = db.detach().numpy()
        out[f'{tag}_v'] = v.numpy()
        out[f'{tag}_ddx'] = (ddx if ddx is not None else torch.zeros_like(x)).detach().numpy()
    # no-bias, dim=1 on a 2-D tensor (mapping FC) and dim=0 / last-dim biases
    x2 = rnd(g, 4, 9)
    b2 = rnd(g, 9)
    out['x2'] = x2.numpy(); out['b2'] = b2.numpy()
    out['y2'] = ref_bias_act.bias_act(x2, b2, dim=1, act='lrelu', impl='ref').numpy()
    out['y2_nob'] = ref_bias_act.bias_act(x2, None, act='relu', gain=1.0, impl='ref').numpy()
    b3 = rnd(g, 7)
    out['b3'] = b3.numpy()
    out['y3_lastdim'] = ref_bias_act.bias_act(x, b3, dim=3, act='swish', impl='ref').numpy()
    np.savez_compressed(os.path.join(OUT, 'bias_act.npz'), **out)
    return len(cases)


UPFIRDN_CASES = [
    # (name, shape, filter taps (1-D list), separable?, up, down, padding, flip, gain)
    ('img_up2',        (2, 3, 8, 8),    [1, 3, 3, 1], None, 2, 1, [2, 1, 2, 1], False, 4),
    ('blur_after_tc',  (2, 4, 17, 17),  [1, 3, 3, 1], None, 1, 1, [1, 1, 1, 1], False, 4),
    ('blur_before_sc', (1, 4, 16, 16),  [1, 3, 3, 1], None, 1, 1, [2, 2, 2, 2], False, 1),
    ('skip_down2',     (2, 4, 16, 16),  [1, 3, 3, 1], None, 1, 2, [1, 1, 1, 1], False, 1),
    ('asym_flip',      (1, 2, 9, 11),   [1, 2, 4, 3, 1], False, [2, 1], [1, 3], [3, 0, 1, 2], True, 2.5),
    ('crop_negpad',    (1, 2, 12, 10),  [1, 3, 3, 1], None, 1, 1, [-1, 2, 0, -2], False, 1),
    ('sep12_up2',      (1, 3, 10, 12),  'sym6', None, 2, 1, [6, 5, 6, 5], False, 4),
    ('sep12_down2',    (1, 3, 20, 24),  'sym6', None, 1, 2, [5, 5, 5, 5], True, 1),
    ('identity_none',  (1, 2, 5, 6),    None, None, 1, 1, 0, False, 1),
    ('up3_down2',      (1, 2, 7, 6),    [1, 4, 6, 4, 1], False, 3, 2, [2, 3, 1, 0], False, 9),
]
SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466,
        0.787641141030194, 0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
        0.0017677118642428036, -0.007800708325034148]   # published sym6 wavelet taps; the filter training/augment.py:35,167 feeds to upfirdn2d


def golden_upfirdn2d(ref_up):
    out = {}
    g = torch.Generator().manual_seed(4321)
    for name, shape, taps, sep, up, down, pad, flip, gain in UPFIRDN_CASES:
        x = rnd(g, *shape)
        if taps is None:
            f = None
        else:
            f = ref_up.setup_filter(SYM6 if taps == 'sym6' else taps, separable=sep)
        xi = x.clone().requires_grad_(True)
        y = ref_up.upfirdn2d(xi, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain, impl='ref')
        dy = rnd(g, *y.shape)
        dx, = torch.autograd.grad(y, [xi], dy)
        out[f'{name}_x'] = x.numpy()
        out[f'{name}_f'] = f.numpy() if f is not None else np.zeros(0, np.float32)
        out[f'{name}_y'] = y.detach().numpy()
        out[f'{name}_dy'] = dy.numpy()
        out[f'{name}_dx'] = dx.numpy()
    # convenience wrappers
    f4 = ref_up.setup_filter([1, 3, 3, 1])
    x = rnd(g, 1, 3, 8, 8)
    out['wrap_x'] = x.numpy()
    out['wrap_f'] = f4.numpy()
    out['wrap_filter2d'] = ref_up.filter2d(x, f4, impl='ref').numpy()
    out['wrap_upsample2d'] = ref_up.upsample2d(x, f4, impl='ref').numpy()
    out['wrap_downsample2d'] = ref_up.downsample2d(x, f4, impl='ref').numpy()
    out['setup_1331'] = f4.numpy()
    out['setup_sym6'] = ref_up.setup_filter(SYM6).numpy()
    out['setup_flip_gain'] = ref_up.setup_filter([1, 2, 3], flip_filter=True, gain=3.0).numpy()
    np.savez_compressed(os.path.join(OUT, 'upfirdn2d.npz'), **out)
    return len(UPFIRDN_CASES)


CONV_CASES = [
    # (name, x shape, w shape, up, down, padding, groups, flip_weight, use filter)
    ('plain3',      (2, 6, 9, 9),   (8, 6, 3, 3), 1, 1, 1, 1, True, False),
    ('plain1',      (2, 6, 9, 9),   (5, 6, 1, 1), 1, 1, 0, 1, True, False),
    ('k7',          (1, 3, 12, 12), (4, 3, 7, 7), 1, 1, 3, 1, True, False),
    ('up2_k3',      (2, 6, 8, 8),   (8, 6, 3, 3), 2, 1, 1, 1, False, True),
    ('up2_k3_grp',  (1, 12, 8, 8),  (8, 6, 3, 3), 2, 1, 1, 2, False, True),
    ('down2_k3',    (2, 6, 16, 16), (8, 6, 3, 3), 1, 2, 1, 1, True, True),
    ('down2_k1',    (2, 6, 16, 16), (8, 6, 1, 1), 1, 2, 0, 1, True, True),
    ('up2_k1',      (2, 6, 8, 8),   (8, 6, 1, 1), 2, 1, 0, 1, True, True),
    ('asym_pad',    (1, 4, 9, 9),   (4, 4, 3, 3), 1, 1, [1, 0, 2, 1], 1, True, False),
    ('up2_down2',   (1, 4, 8, 8),   (4, 4, 3, 3), 2, 2, 1, 1, True, True),
]


def golden_conv(ref_cr, ref_up):
    out = {}
    g = torch.Generator().manual_seed(2468)
    f = ref_up.setup_filter([1, 3, 3, 1])
    out['f'] = f.numpy()
    for name, xs, ws, up, down, pad, groups, flipw, usef in CONV_CASES:
        x = rnd(g, *xs)
        w = rnd(g, *ws, scale=0.3)
        y = ref_cr.conv2d_resample(x, w, f=(f if usef else None), up=up, down=down, padding=pad, groups=groups,
                                   flip_weight=flipw)
        out[f'{name}_x'] = x.numpy(); out[f'{name}_w'] = w.numpy(); out[f'{name}_y'] = y.numpy()
    np.savez_compressed(os.path.join(OUT, 'conv2d_resample.npz'), **out)
    return len(CONV_CASES)


MODCONV_CASES = [
    # (name, N, I, O, k, H, up, demodulate, noise kind, flip_weight)
    ('k3',        2, 16, 24, 3, 8,  1, True,  'hw',   True),
    ('k3_up2',    2, 16, 24, 3, 8,  2, True,  'hw',   False),
    ('k3_nonoise', 3, 8, 8,  3, 6,  1, True,  None,   True),
    ('torgb',     2, 16, 3,  1, 8,  1, False, None,   True),
    ('torgb7',    1, 16, 7,  1, 8,  1, False, None,   True),
    ('k3_n1hw',   2, 8, 8,   3, 8,  2, True,  'n1hw', False),
]


def golden_modconv(networks, ref_up, ref_bias_act):
    out = {}
    g = torch.Generator().manual_seed(1357)
    f = ref_up.setup_filter([1, 3, 3, 1])
    out['f'] = f.numpy()
    for name, n, ic, oc, k, h, up, demod, noise_kind, flipw in MODCONV_CASES:
        x = rnd(g, n, ic, h, h)
        w = rnd(g, oc, ic, k, k)
        s = rnd(g, n, ic, scale=0.5) + 1.0
        ho = h * up
        noise = None
        if noise_kind == 'hw':
            noise = rnd(g, ho, ho, scale=0.1)
        elif noise_kind == 'n1hw':
            noise = rnd(g, n, 1, ho, ho, scale=0.1)
        kw = dict(noise=noise, up=up, padding=k // 2, resample_filter=f, demodulate=demod, flip_weight=flipw)
        y_fused = networks.modulated_conv2d(x.clone(), w, s, fused_modconv=True, **kw)
        y_split = networks.modulated_conv2d(x.clone(), w, s, fused_modconv=False, **kw)
        out[f'{name}_x'] = x.numpy(); out[f'{name}_w'] = w.numpy(); out[f'{name}_s'] = s.numpy()
        out[f'{name}_noise'] = noise.numpy() if noise is not None else np.zeros(0, np.float32)
        out[f'{name}_y_fused'] = y_fused.numpy(); out[f'{name}_y_split'] = y_split.numpy()
        # the layer-level composition the callers use (SynthesisLayer / ToRGB)
        b = rnd(g, oc)
        out[f'{name}_b'] = b.numpy()
        if demod:
            ya = ref_bias_act.bias_act(y_fused, b, act='lrelu', gain=float(np.sqrt(2)), clamp=256.0, impl='ref')
        else:
            ya = ref_bias_act.bias_act(y_fused, b, clamp=256.0, impl='ref')
        out[f'{name}_y_act'] = ya.numpy()
    np.savez_compressed(os.path.join(OUT, 'modulated_conv2d.npz'), **out)
    return len(MODCONV_CASES)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)    # fixed summation order for the library convolutions
    with reference_imports():
        from torch_utils.ops import bias_act as ref_bias_act
        from torch_utils.ops import upfirdn2d as ref_up
        from torch_utils.ops import conv2d_resample as ref_cr
        import training.networks as networks
        n1 = golden_bias_act(ref_bias_act)
        n2 = golden_upfirdn2d(ref_up)
        n3 = golden_conv(ref_cr, ref_up)
        n4 = golden_modconv(networks, ref_up, ref_bias_act)
    print(f'golden fixtures written to {os.path.normpath(OUT)}: bias_act {n1} cases, upfirdn2d {n2}, '
          f'conv2d_resample {n3}, modulated_conv2d {n4}')


if __name__ == '__main__':
    main()
