"""Mint the augmentation fixture from the REAL reference (TEST INFRASTRUCTURE ONLY; build container only).

    python oracle/make_golden_augment.py        # writes tests/golden/augment.npz

Runs /root/reference/training/augment.py's AugmentPipe on the CPU (its ops take their PyTorch reference path there) for a few
configurations, with fixed seeds (random mode) and with `debug_percentile` (deterministic mode, device independent), and stores
inputs and outputs.  The package's AugmentPipe must reproduce them: same draws in the same order -> same images.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle.make_golden import OUT, reference_imports

BGC = dict(xflip=1, rotate90=1, xint=1, scale=1, rotate=1, aniso=1, xfrac=1, brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1)
CONFIGS = {
    'bgc': BGC,                                                                 # train.py:253-260 default ('ada' with augpipe 'bgc')
    'bgcfnc': dict(BGC, imgfilter=1, noise=1, cutout=1),
    'geom_only': dict(scale=1, rotate=1, aniso=1, xfrac=1),
    'color_only': dict(brightness=1, contrast=1, lumaflip=1, hue=1, saturation=1),
    'blit_only': dict(xflip=1, rotate90=1, xint=1),
}
RUNS = [  # name, config, p, seed, debug_percentile, shape
    ('bgc_p1', 'bgc', 1.0, 11, None, (4, 3, 40, 48)),
    ('bgc_p06', 'bgc', 0.6, 12, None, (6, 3, 32, 32)),
    ('bgcfnc_p08', 'bgcfnc', 0.8, 13, None, (4, 3, 48, 40)),
    ('geom_p1', 'geom_only', 1.0, 14, None, (3, 3, 33, 47)),
    ('color_p1', 'color_only', 1.0, 15, None, (3, 3, 16, 24)),
    ('color_gray', 'color_only', 1.0, 16, None, (3, 1, 16, 24)),
    ('blit_p1', 'blit_only', 1.0, 17, None, (5, 3, 24, 24)),
    ('bgc_dbg02', 'bgc', 1.0, 18, 0.2, (2, 3, 40, 48)),
    ('bgc_dbg07', 'bgc', 1.0, 19, 0.7, (2, 3, 64, 64)),
    ('bgcfnc_dbg09', 'bgcfnc', 1.0, 20, 0.9, (2, 3, 32, 40)),
]


def main():
    torch.set_num_threads(4)
    out = {}
    with reference_imports():
        import training.augment as augment
        for name, cfg, p, seed, dbg, shape in RUNS:
            pipe = augment.AugmentPipe(**CONFIGS[cfg]).eval().requires_grad_(False)
            pipe.p.copy_(torch.as_tensor(p))
            x = torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).clamp(-1, 1)
            torch.manual_seed(seed + 100)
            y = pipe(x, debug_percentile=dbg)
            out[f'{name}/x'] = x.numpy()
            out[f'{name}/y'] = y.numpy()
        out['Hz_geom'] = pipe.Hz_geom.numpy()
        out['Hz_fbank'] = pipe.Hz_fbank.numpy()
    np.savez_compressed(os.path.join(OUT, 'augment.npz'), **out)
    print('wrote augment.npz', {k: v.shape for k, v in out.items() if k.endswith('/y')})


if __name__ == '__main__':
    main()
