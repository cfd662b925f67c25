"""BASELINE configs[0] input: the real try-on pairs of tests/golden/test_pair_upper.npz, written by the REFERENCE's own loader
(oracle/make_golden_testpair.py over /root/reference/test_datas).  CPU-side checks of the fixture and of the host half of the input
edge (test.py:121-147) on it."""
import importlib
import os

import numpy as np
import torch

from conftest import GOLDEN, load_pkg
from oracle import ref_io

load_pkg()
tryon_io = importlib.import_module('pgpp_b200.tryon_io')
FIX = os.path.join(GOLDEN, 'test_pair_upper.npz')


def test_fixture_has_the_dataset_tuple_of_test_py():
    d = tryon_io.load_test_pairs(FIX, (0, 1))
    want = {'image': 3, 'clothes': 3, 'pose': 3, 'clothes_pose': 3, 'norm_img': 30, 'norm_img_lower': 15, 'denorm_upper_clothes': 3,
            'denorm_lower_clothes': 3, 'denorm_upper_mask': 1, 'denorm_lower_mask': 1, 'retain_mask': 1, 'skin_average': 3,
            'lower_label_map': 1, 'lower_clothes_upper_bound': 1}
    for k, c in want.items():
        hw = (128, 128) if k.startswith('norm_img') else (512, 512)
        assert tuple(d[k].shape) == (2, c) + hw, (k, d[k].shape)
    assert len(d['person_name']) == 2 and d['person_name'][0].endswith('.jpg')
    # 512 x 320 photographs padded to 512 x 512 with white (dataset.py:2038-2046): the 96-pixel side bands are constant
    img = d['image']
    assert img.dtype == torch.uint8 and int(img[..., :96].min()) == 255 and int(img[..., 416:].min()) == 255
    assert int(d['pose'][..., :96].max()) == 0
    assert set(np.unique(d['retain_mask'].numpy())) <= {0, 1}
    assert set(np.unique(d['denorm_upper_mask'].numpy())) <= {0, 1}


def test_reference_input_edge_on_the_real_pair():
    d = tryon_io.load_test_pairs(FIX, (0, 1))
    x = ref_io.prepare_inputs(d)
    assert tuple(x['parts'].shape) == (2, 45, 128, 128) and tuple(x['pose'].shape) == (2, 5, 512, 512) and tuple(x['retain'].shape) == (2, 6, 512, 512)
    for k in ('image', 'parts', 'pose', 'retain', 'denorm_upper_clothes', 'denorm_lower_clothes'):
        assert float(x[k].min()) >= -1.0 and float(x[k].max()) <= 1.0
    # lower_label_map is 0 / 127.5 / 255 (trousers / skirt / dress, dataset.py:2212-2219) -> -1 / 0 / +1
    assert set(np.unique(x['pose'][:, 3].numpy())) <= {-1.0, 0.0, 1.0}
