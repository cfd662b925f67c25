"""Mint the checkpoint-format fixture from the REAL reference (TEST INFRASTRUCTURE ONLY; build container only).

    python oracle/make_golden_snapshot.py        # writes tests/golden/ref_snapshot_small.pkl + ref_snapshot_small.npz

Pickles small networks of /root/reference/training/networks.py exactly the way the training loop writes
`network-snapshot-*.pkl` (training_loop_fullbody.py:723-736: `pickle.dump(dict(G=..., D=..., G_ema=...,
training_set_kwargs=..., augment_pipe=...))` of persistent-class modules, torch_utils/persistence.py:118-126), together with
their state dicts and a reference forward on seeded inputs.  The loader under test (pasta-gan-plusplus_b200/legacy.py) must
read the file WITHOUT the reference tree being importable.

It then pickles the full-size GeneratorFull_v20 to a temporary file, reads it back through the loader and checks that all
338 tensors load strictly into this package's generator (printed; too large to commit).
"""
import io
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from oracle import ref_generator
from oracle.make_golden import OUT, reference_imports
from oracle import make_golden_generator as mgg


def main():
    torch.set_num_threads(8)
    with reference_imports():
        import dnnlib
        from torch_utils.ops import bias_act, upfirdn2d
        import training.networks as networks
        mgg._REF.update(networks=networks, upfirdn2d=upfirdn2d, bias_act=bias_act)
        networks.SynthesisLayer = mgg.SynthesisLayer
        torch.manual_seed(0)
        nets = dict(
            G=networks.MappingNetwork(z_dim=0, c_dim=32, w_dim=32, num_ws=6, num_layers=1),
            D=networks.ResBlock(8, 16, 3, down=2),
            G_ema=networks.StyleEncoderNetworkV18(input_nc=5, output_nc=32, ngf=4),
        )
        for net in nets.values():
            ref_generator.name_seeded_init(list(net.named_parameters()) + list(net.named_buffers()))
            net.eval().requires_grad_(False)
        snapshot = dict(training_set_kwargs=dict(dnnlib.EasyDict(class_name='training.dataset.UvSamplerPartsDataset', path='data.zip',
                                                                 use_labels=True, max_size=np.int64(100), xflip=False, resolution=512)))
        snapshot.update(nets)
        snapshot['augment_pipe'] = None
        buf = io.BytesIO()
        pickle.dump(snapshot, buf)
        with open(os.path.join(OUT, 'ref_snapshot_small.pkl'), 'wb') as f:
            f.write(buf.getvalue())
        g = torch.Generator().manual_seed(5)
        out = {}
        c = torch.randn(3, 32, generator=g)
        x8 = torch.randn(2, 8, 16, 16, generator=g)
        x5 = torch.randn(2, 5, 64, 64, generator=g)
        x6 = torch.randn(2, 6, 64, 64, generator=g)
        with torch.no_grad():
            out['G_in_c'] = c.numpy(); out['G_out'] = nets['G'](torch.zeros(3, 0), c).numpy()
            out['D_in'] = x8.numpy(); out['D_out'] = nets['D'](x8).numpy()
            style, feats = nets['G_ema'](x5, x6)
            out['E_in_x'] = x5.numpy(); out['E_in_const'] = x6.numpy(); out['E_out_style'] = style.numpy()
            for i, ft in enumerate(feats):
                out[f'E_out_feat{i}'] = ft.numpy()
        for key, net in nets.items():
            for name, v in net.state_dict().items():
                out[f'sd/{key}/{name}'] = v.numpy()
        np.savez_compressed(os.path.join(OUT, 'ref_snapshot_small.npz'), **out)
        print('wrote ref_snapshot_small.pkl', len(buf.getvalue()), 'bytes')

        # full-size generator round trip (not committed)
        G = networks.GeneratorFull_v20(z_dim=0, c_dim=512, w_dim=512, img_resolution=512, img_channels=3, mapping_kwargs=dict(num_layers=1),
                                       synthesis_kwargs=dict(channel_base=32768, channel_max=512, num_fp16_res=3, conv_clamp=256,
                                                             use_noise=True)).eval().requires_grad_(False)
        ref_sd = {k: v.clone() for k, v in G.state_dict().items()}
        tmp = tempfile.NamedTemporaryFile(suffix='.pkl', delete=False)
        pickle.dump(dict(G=G, D=nets['D'], G_ema=G, training_set_kwargs=None, augment_pipe=None), tmp)
        tmp.close()
    # the reference is no longer importable from here on
    for name in [m for m in sys.modules if m.split('.')[0] in ('training', 'torch_utils', 'dnnlib', 'legacy')]:
        del sys.modules[name]
    from __graft_entry__ import load_pkg
    load_pkg()
    from pgpp_b200 import legacy
    with open(tmp.name, 'rb') as f:
        data = legacy.load_network_pkl(f, plain_modules=('oracle.make_golden_generator',))   # the SynthesisLayer shim
    os.unlink(tmp.name)
    rec = data['G_ema']
    sd = legacy.state_dict(rec)
    assert set(sd) == set(ref_sd) and all(torch.equal(sd[k], ref_sd[k]) for k in sd)
    ours = legacy.build_generator(rec)
    print('full-size snapshot:', rec, '-> strict load into', type(ours).__name__, 'ok;',
          'embedded source readable:', len(legacy.class_source(rec, 'GeneratorFull_v20')), 'chars of class GeneratorFull_v20')


if __name__ == '__main__':
    main()
