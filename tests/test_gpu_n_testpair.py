"""BASELINE configs[0] on the GPU: a REAL upper-body try-on pair from the reference's test_datas (fixture written by the reference's
own loader, oracle/make_golden_testpair.py) through the input edge, the full generator and the output edge of this package, against
the CPU restatement of the reference path on the same tensors and the same (name-seeded random) weights."""
import importlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from oracle import ref_generator, ref_io

pytestmark = pytest.mark.gpu
load_pkg()
gen = importlib.import_module('pgpp_b200.training.generator')
tryon_io = importlib.import_module('pgpp_b200.tryon_io')
cg = importlib.import_module('pgpp_b200.torch_utils.ops.conv2d_gradfix')
FIX = os.path.join(GOLDEN, 'test_pair_upper.npz')
DEV = 'cuda:0'


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope='module')
def net():
    G = gen.build_generator().eval()
    ref_generator.name_seeded_init(list(G.named_parameters()) + list(G.named_buffers()))
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    return G.to(DEV).requires_grad_(False), sd


def test_input_edge_on_real_pairs_is_bit_identical():
    data = tryon_io.load_test_pairs(FIX, (0, 1))
    want = ref_io.prepare_inputs(data)
    got = tryon_io.prepare_inputs(data, DEV)
    for k in want:
        assert torch.equal(got[k].cpu(), want[k]), k


def test_tryon_on_a_real_pair_matches_the_cpu_reference_path(net):
    G, sd = net
    data = tryon_io.load_test_pairs(FIX, (0,))
    x = tryon_io.prepare_inputs(data, DEV)
    xc = ref_io.prepare_inputs(data)
    with torch.no_grad():
        # test.py passes gt_parsing=None: the parsing the synthesis network predicts decides the upper / lower masks
        r_img, r_fin, r_par = ref_generator.generator(sd, xc['parts'], xc['retain'], xc['pose'], xc['denorm_upper_clothes'], xc['denorm_lower_clothes'],
                                                      xc['denorm_upper_mask'], xc['denorm_lower_mask'], None)
    img, fin, par = tryon_io.tryon(G, x, noise_mode='const')
    assert rel(img, r_img) < 1e-4 and rel(par, r_par) < 1e-4, (rel(img, r_img), rel(par, r_par))
    # the discrete decision (argmax over the 7 parsing classes) agrees except where two logits tie to rounding error
    a, b = par.argmax(dim=1).cpu(), r_par.argmax(dim=1)
    flips = int((a != b).sum())
    assert flips <= 8, flips
    if flips == 0:
        assert rel(fin, r_fin) < 1e-4, rel(fin, r_fin)
        assert float((fin.cpu() - r_fin).abs().max()) <= 1e-3 * max(1.0, float(r_fin.abs().max()))
    # with the parsing fixed to the reference's decision nothing discrete depends on rounding: the north-star bars hold
    gt = b[:, None].float()
    with torch.no_grad():
        _, r_fin2, _ = ref_generator.generator(sd, xc['parts'], xc['retain'], xc['pose'], xc['denorm_upper_clothes'], xc['denorm_lower_clothes'],
                                               xc['denorm_upper_mask'], xc['denorm_lower_mask'], gt)
    _, fin2, _ = tryon_io.tryon(G, x, gt_parsing=gt.to(DEV), noise_mode='const')
    assert rel(fin2, r_fin2) < 1e-4, rel(fin2, r_fin2)
    assert float((fin2.cpu() - r_fin2).abs().max()) <= 1e-3 * max(1.0, float(r_fin2.abs().max()))
    # output edge (test.py:162-166) of both: identical uint8 images except where a value sits on a rounding boundary
    u8 = tryon_io.images_to_uint8(fin2).cpu().numpy()
    want = ref_io.images_to_uint8(r_fin2)
    assert u8.shape == want.shape == (1, 512, 512, 3)
    assert float((u8.astype(np.int32) - want.astype(np.int32)).__abs__().max()) <= 1
    assert float((u8 != want).mean()) < 1e-3
