"""TEST INFRASTRUCTURE.  Writes tests/golden/test_pair_upper.npz: the uint8 tensors the REAL reference loader
(/root/reference/training/dataset.py:1952-2223, UvitonDatasetFull_512_test_upper, exactly as test.py:112-124 uses it) produces for the
first pairs of /root/reference/test_datas/test_pairs.txt - BASELINE configs[0]: "one upper-body pair from test_datas".

The loader is imported unmodified; three of its third-party imports are absent from this image and are shimmed:
  skimage.draw.circle     (scikit-image 0.18.3, README.md:16; removed in >= 0.19): pixels with (r - r0)^2 + (c - c0)^2 < radius^2, clipped
                          to `shape` - the published definition of the function
  skimage.draw.line_aa    imported by dataset.py:20 but never called on the test path
  pycocotools.mask        frPyObjects / merge / decode of ONE polygon (dataset.py:2243-2248, the palm rectangles): rasterised here with
                          cv2.fillPoly; pixels exactly on a polygon edge may differ from pycocotools' 5x-supersampled boundary walk
so the fixture is a faithful real try-on pair whose keypoint discs / palm masks can differ from a run with the pinned packages in a
few boundary pixels.  It is an INPUT fixture: parity is always between this repo's generator and the reference path on these same
tensors.
    python oracle/make_golden_testpair.py [n_pairs=2]
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
KEYS = ('image', 'clothes', 'pose', 'clothes_pose', 'norm_img', 'norm_img_lower', 'denorm_upper_img', 'denorm_lower_img', 'denorm_upper_mask',
        'denorm_lower_mask', 'retain_mask', 'skin_average', 'lower_label_map', 'lower_clothes_upper_bound')


def _circle(r, c, radius, shape=None):
    rad = int(np.ceil(radius))
    rr, cc = np.mgrid[r - rad:r + rad + 1, c - rad:c + rad + 1]
    keep = (rr - r) ** 2 + (cc - c) ** 2 < radius ** 2
    rr, cc = rr[keep], cc[keep]
    if shape is not None:
        ok = (rr >= 0) & (rr < shape[0]) & (cc >= 0) & (cc < shape[1])
        rr, cc = rr[ok], cc[ok]
    return rr, cc


def _install_shims():
    import cv2
    for m in ('matplotlib', 'matplotlib.pyplot', 'skimage', 'skimage.draw', 'pycocotools', 'pycocotools.mask'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['skimage.draw'].circle = _circle
    sys.modules['skimage.draw'].line_aa = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError('line_aa is not on the test path'))
    mu = sys.modules['pycocotools.mask']
    mu.frPyObjects = lambda polys, h, w: [(np.asarray(p, np.float64).reshape(-1, 2), h, w) for p in polys]
    mu.merge = lambda rles: rles

    def decode(rles):
        h, w = rles[0][1], rles[0][2]
        m = np.zeros((h, w), np.uint8)
        for pts, _, _ in rles:
            cv2.fillPoly(m, [np.round(pts).astype(np.int32)], 1)
        return m
    mu.decode = decode
    sys.modules['pycocotools'].mask = mu


def main(n_pairs=2):
    _install_shims()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import training.dataset as ds
        data = ds.UvitonDatasetFull_512_test_upper(path=os.path.join(REF, 'test_datas'), test_txt='test_pairs.txt', use_sleeve_mask=False,
                                                   max_size=None, xflip=False)
        out = {}
        names = []
        for i in range(n_pairs):
            item = data[i]
            assert len(item) == 16, len(item)
            for k, v in zip(KEYS, item[:14]):
                v = np.asarray(v)
                out[f'{k}_{i}'] = v.astype(np.uint8) if v.dtype != np.uint8 and float(np.abs(v - np.round(v)).max()) == 0 and v.min() >= 0 and v.max() <= 255 else v
            names.append((str(item[14]), str(item[15])))
        out['names'] = np.array(names)
    finally:
        os.chdir(cwd)
    path = os.path.join(ROOT, 'tests', 'golden', 'test_pair_upper.npz')
    np.savez_compressed(path, **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype, (int(v.min()), int(v.max())) if v.dtype != np.dtype('<U') and v.dtype.kind != 'U' else '')
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
