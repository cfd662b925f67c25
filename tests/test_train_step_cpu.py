"""The training-step harness (pgpp_b200.training.training_step / loss, modelled on training_loop_fullbody.py:452-481,603-650 and
loss_fullbody.py:115-330) on the CPU: phase bookkeeping, and a world_size-2 gloo run of the data-parallel `Dboth` phase (real-image
loss + R1 penalty + generated-image loss through the frozen generator) whose DDP-averaged gradients must equal the single-process
step on the concatenated batch.  CPU tensors take the ops' PyTorch reference path (custom_ops.cpu_tensors = 'ref')."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = dict(channel_base=4096, channel_max=512, mbstd_group_size=1, num_fp16_res=0)     # 8 channels at 512 px, fp32 blocks (fp16 rounding is batch-size dependent)


def _setup():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from conftest import load_pkg
    load_pkg()
    importlib.import_module('pgpp_b200.torch_utils.custom_ops').cpu_tensors = 'ref'
    return importlib.import_module('pgpp_b200.training.training_step')


def _build(ts, distributed):
    torch.manual_seed(0)
    G, D, DP = ts.build_networks('cpu', **SMALL)
    return ts.TrainingStep(G, D, DP, 'cpu', batch_size=2, distributed=distributed)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(4)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ts = _setup()
    step = _build(ts, distributed=True)
    assert isinstance(step.ddp_modules['D'], torch.nn.parallel.DistributedDataParallel)
    data = {k: v[rank:rank + 1] for k, v in ts.synthetic_batch(2, 'cpu', seed=3).items()}
    stats = step(data, phases=['Dboth'])
    if rank == 0:
        torch.save({'grads': {n: p.grad.clone() for n, p in step.D.named_parameters()}, 'stats': stats}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dboth_phase_equals_single_process(tmp_path):
    ts = _setup()
    out_path = str(tmp_path / 'grads.pt')
    port = 35500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    torch.set_num_threads(8)
    step = _build(ts, distributed=False)
    assert [p['name'] for p in step.phases] == ['Gboth', 'Dboth', 'D_parsingboth', 'D_parsingboth']
    w_before = {n: p.detach().clone() for n, p in step.D.named_parameters()}
    step(ts.synthetic_batch(2, 'cpu', seed=3), phases=['Dboth'])
    got = torch.load(out_path)
    for n, p in step.D.named_parameters():
        scale = max(p.grad.abs().max().item(), 1e-8)
        assert (got['grads'][n] - p.grad).abs().max().item() <= 2e-4 * scale, n
    assert 'Loss/r1_penalty' in got['stats'] and step.batch_idx == 1
    assert any(not torch.equal(w_before[n], p) for n, p in step.D.named_parameters())        # Adam moved the weights
    assert all(not p.requires_grad for p in step.D.parameters())                               # phases leave the modules frozen


def test_lazy_regularisation_phase_list():
    ts = _setup()
    torch.manual_seed(0)
    G, D, DP = ts.build_networks('cpu', **SMALL)
    step = ts.TrainingStep(G, D, DP, 'cpu', G_reg_interval=4, D_reg_interval=16, distributed=False)
    assert [(p['name'], p['interval']) for p in step.phases] == [
        ('Gmain', 1), ('Greg', 4), ('Dmain', 1), ('Dreg', 16), ('D_parsingmain', 1), ('D_parsingreg', 16), ('D_parsingmain', 1), ('D_parsingreg', 16)]
    lr = step.phases[2]['opt'].param_groups[0]
    assert abs(lr['lr'] - 0.002 * 16 / 17) < 1e-9 and abs(lr['betas'][1] - 0.99 ** (16 / 17)) < 1e-9
