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


# ---- training step (BASELINE config 4): data-parallel discriminator step with the R1 penalty, gradients all-reduced -------------

def _d_loss(D, img):
    img = img.detach().requires_grad_(True)
    logits = D(img, None, fused=False, impl='ref')
    gx, = torch.autograd.grad(logits.sum(), [img], create_graph=True)
    return (torch.nn.functional.softplus(-logits).squeeze(1) + 5.0 * gx.square().sum([1, 2, 3])).mean()


def _build_d():
    disc = importlib.import_module('pgpp_b200.training.discriminator')
    torch.manual_seed(0)
    D = disc.Discriminator(c_dim=0, img_resolution=16, img_channels=3, channel_base=128, channel_max=16,
                           epilogue_kwargs=dict(mbstd_group_size=2)).train()
    imgs = torch.randn(4, 3, 16, 16, generator=torch.Generator().manual_seed(2)).clamp(-1, 1)
    return D, imgs


def _train_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    _, upfirdn2d = _setup()
    from helpers import upfirdn2d_ref_on_cpu
    D, imgs = _build_d()
    per = imgs.shape[0] // world
    with upfirdn2d_ref_on_cpu(upfirdn2d):
        _d_loss(D, imgs[rank * per:(rank + 1) * per]).backward()
    # the exchange step of the training path: gradient all-reduce (NCCL over NVLink on the GPUs, gloo here), averaged
    for p in D.parameters():
        dist.all_reduce(p.grad)
        p.grad /= world
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in D.named_parameters()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_d_step_with_r1_equals_single_process(tmp_path):
    _, upfirdn2d = _setup()
    from helpers import upfirdn2d_ref_on_cpu
    out_path = str(tmp_path / 'grads.pt')
    port = 31500 + os.getpid() % 2000
    mp.spawn(_train_worker, args=(2, port, out_path), nprocs=2, join=True)
    D, imgs = _build_d()
    with upfirdn2d_ref_on_cpu(upfirdn2d):
        # minibatch-std groups of 2 stay inside a rank's shard, so the mean of the two shard losses is the global loss
        (0.5 * (_d_loss(D, imgs[:2]) + _d_loss(D, imgs[2:]))).backward()
    got = torch.load(out_path)
    for n, p in D.named_parameters():
        assert torch.allclose(got[n], p.grad, rtol=1e-4, atol=1e-6), n
