"""world_size-2 gloo test (CPU) of the batch-sharded inference path: every rank runs its shard of a global batch through
the synthesis chain (composition route, impl='ref'), the shards are gathered and must equal the single-process result.
No collective is used on the data path itself; gloo is only the test's transport for the comparison."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from conftest import load_pkg
    load_pkg()
    return (importlib.import_module('pgpp_b200.training.synthesis'), importlib.import_module('pgpp_b200.torch_utils.ops.upfirdn2d'))


def _build(synthesis):
    torch.manual_seed(0)
    net = synthesis.SynthesisChain(w_dim=32, img_resolution=32, channel_base=512, channel_max=16, merge_channels=0).eval()
    g = torch.Generator().manual_seed(1)
    ws = torch.randn(5, net.num_ws, 32, generator=g)
    pose = torch.randn(5, net.channels[8], 8, 8, generator=g)
    return net, ws, pose


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    synthesis, upfirdn2d = _setup()
    from helpers import upfirdn2d_ref_on_cpu
    net, ws, pose = _build(synthesis)
    with torch.no_grad(), upfirdn2d_ref_on_cpu(upfirdn2d):
        (a, b), (img, parsing, tex) = synthesis.run_sharded(net, ws, pose, None, rank, world, fused=False, impl='ref', noise_mode='const')
    full = torch.zeros(5, 3, 32, 32)
    full[a:b] = img
    dist.all_reduce(full)           # disjoint shards: the sum assembles the global batch
    if rank == 0:
        torch.save({'img': full, 'ranges': (a, b)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_inference_equals_single_process(tmp_path):
    synthesis, upfirdn2d = _setup()
    from helpers import upfirdn2d_ref_on_cpu
    assert [synthesis.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [synthesis.shard_range(256, r, 8) for r in range(8)][-1] == (224, 256)
    out_path = str(tmp_path / 'gathered.pt')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    net, ws, pose = _build(synthesis)
    with torch.no_grad(), upfirdn2d_ref_on_cpu(upfirdn2d):
        want, _, _ = net(ws, pose, None, fused=False, impl='ref', noise_mode='const')
    got = torch.load(out_path)['img']
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
