"""Checkpoint format: `pgpp_b200.legacy.load_network_pkl` reads a snapshot written by the REAL reference
(tests/golden/ref_snapshot_small.pkl, minted by oracle/make_golden_snapshot.py through persistence.persistent_class +
pickle.dump, as training_loop_fullbody.py:723-736 does) without importing or executing anything from the file, and the
networks rebuilt from it reproduce the reference's forward outputs."""
import io
import os
import pickle
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_pkg
from helpers import upfirdn2d_ref_on_cpu

load_pkg()
from pgpp_b200 import legacy
from pgpp_b200.torch_utils.ops import upfirdn2d as _upfirdn2d

PKL = os.path.join(GOLDEN, 'ref_snapshot_small.pkl')
NPZ = os.path.join(GOLDEN, 'ref_snapshot_small.npz')


@pytest.fixture(scope='module')
def snapshot():
    assert not any(m in sys.modules for m in ('torch_utils.persistence', 'dnnlib', 'training.networks')), 'reference must not be importable'
    with open(PKL, 'rb') as f:
        return legacy.load_network_pkl(f)


@pytest.fixture(scope='module')
def golden():
    return dict(np.load(NPZ))


def test_snapshot_structure_and_tensors(snapshot, golden):
    assert set(snapshot) == {'G', 'D', 'G_ema', 'training_set_kwargs', 'augment_pipe'}
    assert snapshot['augment_pipe'] is None
    ts = snapshot['training_set_kwargs']
    assert ts['class_name'] == 'training.dataset.UvSamplerPartsDataset' and ts['resolution'] == 512 and int(ts['max_size']) == 100
    assert [snapshot[k].class_name for k in ('G', 'D', 'G_ema')] == ['MappingNetwork', 'ResBlock', 'StyleEncoderNetworkV18']
    for key in ('G', 'D', 'G_ema'):
        sd = legacy.state_dict(snapshot[key])
        want = {k[len(f'sd/{key}/'):]: v for k, v in golden.items() if k.startswith(f'sd/{key}/')}
        assert set(sd) == set(want)
        for name, v in sd.items():
            assert v.dtype == torch.float32 and np.array_equal(v.numpy(), want[name]), (key, name)
    assert snapshot['G'].init_kwargs == dict(z_dim=0, c_dim=32, w_dim=32, num_ws=6, num_layers=1)
    assert snapshot['D'].init_args == (8, 16, 3) and snapshot['D'].init_kwargs == dict(down=2)
    tree = legacy.module_tree(snapshot['G_ema'])
    assert tree[''] == 'StyleEncoderNetworkV18' and tree['model'] == 'Sequential' and tree['model.0'] == 'Conv2dLayer'
    assert tree['model.1'] == 'Dense' and tree['model.1.bn'] == 'InstanceNorm2d' and tree['fc'] == 'FullyConnectedLayer'


def test_rebuilt_networks_reproduce_the_reference_forward(snapshot, golden):
    t = lambda k: torch.from_numpy(golden[k])
    with torch.no_grad(), upfirdn2d_ref_on_cpu(_upfirdn2d):
        G = legacy.build_module(snapshot['G'])
        assert torch.allclose(G(torch.zeros(3, 0), t('G_in_c'), impl='ref'), t('G_out'), rtol=1e-5, atol=1e-5)
        D = legacy.build_module(snapshot['D'])
        assert torch.allclose(D(t('D_in'), fused=False, impl='ref'), t('D_out'), rtol=1e-5, atol=1e-5)
        E = legacy.build_module(snapshot['G_ema'])
        style, feats = E(t('E_in_x'), t('E_in_const'), fused=False, impl='ref')
        assert torch.allclose(style, t('E_out_style'), rtol=1e-4, atol=1e-5)
        assert len(feats) == 4
        for i, ft in enumerate(feats):
            assert torch.allclose(ft, t(f'E_out_feat{i}'), rtol=1e-4, atol=1e-5)
    assert not any(p.requires_grad for p in E.parameters()) and not E.training


def test_embedded_source_is_readable_but_never_executed(snapshot):
    src = legacy.class_source(snapshot['G'], 'MappingNetwork')
    assert src.startswith('class MappingNetwork') and 'w_avg' in src
    with pytest.raises(KeyError):
        legacy.class_source(snapshot['G'], 'NoSuchClass')
    assert not any(m.startswith('_imported_module_') for m in sys.modules)      # the reference's exec() path leaves these behind
    with pytest.raises(ValueError):
        legacy.build_generator(snapshot['G'])                                    # not a GeneratorFull_v20


class _Evil:
    def __reduce__(self):
        return (os.system, ('echo pwned > /dev/null',))


@pytest.mark.parametrize('payload', [_Evil(), dict(G=_Evil()), eval, io.BytesIO], ids=['os.system', 'nested', 'builtins.eval', 'io.BytesIO'])
def test_pickles_that_reach_outside_the_allow_list_are_rejected(payload):
    blob = pickle.dumps(payload)
    with pytest.raises(pickle.UnpicklingError, match='allow-list'):
        legacy.load_network_pkl(io.BytesIO(blob))


def test_non_snapshot_pickles_are_rejected():
    with pytest.raises(pickle.UnpicklingError, match='not a PASTA-GAN'):
        legacy.load_network_pkl(io.BytesIO(pickle.dumps([1, 2, 3])))
    with pytest.raises(pickle.UnpicklingError, match='not a persistent network'):
        legacy.load_network_pkl(io.BytesIO(pickle.dumps(dict(G_ema={'weights': 1}))))
    with pytest.raises(AssertionError):
        legacy.load_network_pkl(io.BytesIO(b''), force_fp16=True)
