"""One iteration of the reference's training loop (`training/training_loop_fullbody.py:452-481,603-650`) around the
hot path: phase list with (optionally lazy) regularisation, per-phase `zero_grad -> requires_grad_(True) ->
accumulate_gradients -> requires_grad_(False) -> nan_to_num -> Adam step`, then the G_ema update.  Data-parallel as in
the reference: `G.mapping`, `G.synthesis`, `G.const_encoding`, `G.style_encoding`, `D`, `D_parsing` are each wrapped in
DistributedDataParallel (`broadcast_buffers=False, find_unused_parameters=True`, :452-460) when a process group is up, so
the gradient all-reduce (NCCL over NVLink on the GPUs) overlaps the backward pass bucket by bucket.

Only what BASELINE configs[4] needs is here: no dataset, snapshot, metric, ADA-controller or logging code.
"""
import copy

import torch

from . import discriminator as _disc
from . import generator as _gen
from .loss import StyleGAN2Loss


def build_networks(device, resolution=512, channel_base=32768, channel_max=512, num_fp16_res=3, conv_clamp=256, mbstd_group_size=4,
                   c_dim=512, w_dim=512, mapping_layers=1):
    """G, D (try-on image + 3 pose channels) and D_parsing (7 parsing classes + 3 pose channels) with the keyword arguments of
    train.py:191-199 and training_loop_fullbody.py:405-410."""
    G = _gen.GeneratorFull_v20(z_dim=0, c_dim=c_dim, w_dim=w_dim, img_resolution=resolution, img_channels=3,
                               mapping_kwargs=dict(num_layers=mapping_layers),
                               synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, num_fp16_res=num_fp16_res,
                                                     conv_clamp=conv_clamp, use_noise=True))
    d_kw = dict(c_dim=c_dim, img_resolution=resolution, channel_base=channel_base, channel_max=channel_max, num_fp16_res=num_fp16_res,
                conv_clamp=conv_clamp, epilogue_kwargs=dict(mbstd_group_size=mbstd_group_size))
    D = _disc.Discriminator(img_channels=3 + 3, **d_kw)
    D_parsing = _disc.Discriminator(img_channels=7 + 3, **d_kw)
    return [m.train().requires_grad_(False).to(device) for m in (G, D, D_parsing)]


class TrainingStep:
    def __init__(self, G, D, D_parsing, device, lr=0.002, betas=(0.0, 0.99), G_reg_interval=None, D_reg_interval=None, r1_gamma=10.0,
                 l1_weight=10.0, mask_weight=30.0, augment_pipe=None, ema_kimg=10.0, batch_size=64, distributed=None, fused_adam=None):
        self.G, self.D, self.D_parsing, self.device = G, D, D_parsing, device
        self.G_ema = copy.deepcopy(G).eval()
        self.batch_size, self.ema_kimg = batch_size, ema_kimg
        if distributed is None:
            distributed = torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
        self.distributed = distributed
        ddp = {}
        for name, module in (('G_mapping', G.mapping), ('G_synthesis', G.synthesis), ('G_const_encoding', G.const_encoding),
                             ('G_style_encoding', G.style_encoding), ('D', D), ('D_parsing', D_parsing)):
            if distributed and len(list(module.parameters())) != 0:
                module.requires_grad_(True)
                ids = [device] if torch.device(device).type == 'cuda' else None
                module = torch.nn.parallel.DistributedDataParallel(module, device_ids=ids, broadcast_buffers=False, find_unused_parameters=True)
                module.requires_grad_(False)
            ddp[name] = module
        self.ddp_modules = ddp
        self.loss = StyleGAN2Loss(device=device, **ddp, augment_pipe=augment_pipe, r1_gamma=r1_gamma, l1_weight=l1_weight, mask_weight=mask_weight)
        if fused_adam is None:
            fused_adam = torch.device(device).type == 'cuda'
        self.phases = []
        # the reference lists D_parsing twice (training_loop_fullbody.py:465-467): two optimisers, two passes per iteration
        for name, module, reg_interval in (('G', G, G_reg_interval), ('D', D, D_reg_interval), ('D_parsing', D_parsing, D_reg_interval),
                                           ('D_parsing', D_parsing, D_reg_interval)):
            if reg_interval is None:
                opt = torch.optim.Adam(module.parameters(), lr=lr, betas=betas, eps=1e-8, fused=fused_adam)
                self.phases.append(dict(name=name + 'both', module=module, opt=opt, interval=1))
            else:       # lazy regularisation (:472-478)
                mb_ratio = reg_interval / (reg_interval + 1)
                opt = torch.optim.Adam(module.parameters(), lr=lr * mb_ratio, betas=[b ** mb_ratio for b in betas], eps=1e-8, fused=fused_adam)
                self.phases.append(dict(name=name + 'main', module=module, opt=opt, interval=1))
                self.phases.append(dict(name=name + 'reg', module=module, opt=opt, interval=reg_interval))
        self.batch_idx = 0
        self.cur_nimg = 0

    def __call__(self, data, phases=None):
        """data: dict with real_img, style_input (45 ch @128), retain, pose, denorm_upper_input, denorm_lower_input,
        denorm_upper_mask, denorm_lower_mask, gt_parsing -- this rank's shard (batch_gpu samples).  One accumulation round
        (batch_size == batch_gpu * world_size), so every phase synchronises its gradients."""
        n = data['real_img'].shape[0]
        stats = {}
        for phase in self.phases:
            if self.batch_idx % phase['interval'] != 0 or (phases is not None and phase['name'] not in phases):
                continue
            gen_z = torch.randn([n, self.G.z_dim], device=self.device)
            phase['opt'].zero_grad(set_to_none=True)
            phase['module'].requires_grad_(True)
            stats.update(self.loss.accumulate_gradients(phase=phase['name'], gen_z=gen_z, sync=True, gain=phase['interval'], **data))
            phase['module'].requires_grad_(False)
            for p in phase['module'].parameters():
                if p.grad is not None:
                    torch.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
            phase['opt'].step()
        # G_ema (:625-632)
        with torch.no_grad():
            ema_nimg = self.ema_kimg * 1000
            beta = 0.5 ** (self.batch_size / max(ema_nimg, 1e-8))
            p_ema, p = list(self.G_ema.parameters()), list(self.G.parameters())
            torch._foreach_lerp_(p_ema, p, 1.0 - beta)          # p_ema = p.lerp(p_ema, beta)
            b_ema, b = list(self.G_ema.buffers()), list(self.G.buffers())
            if b:
                torch._foreach_copy_(b_ema, b)
        self.cur_nimg += self.batch_size
        self.batch_idx += 1
        return stats


def synthetic_batch(n, device, seed=0, resolution=512):
    """synthetic training tensors of SURVEY 8d ("training step"): shapes of training_loop_fullbody.py:423-431 after the data-fetch
    arithmetic of :540-590, 512 x 320 content in columns 96..415."""
    g = torch.Generator().manual_seed(seed)
    r = resolution
    lo, hi = (96 * r) // 512, (416 * r) // 512
    band = torch.zeros(1, 1, 1, r, dtype=torch.bool); band[..., lo:hi] = True

    def img(ch, fill):
        t = torch.randn(n, ch, r, r, generator=g).clamp_(-1, 1)
        return torch.where(band, t, torch.full_like(t, fill))
    d = dict(real_img=img(3, 1.0), style_input=torch.randn(n, 45, r // 4, r // 4, generator=g).clamp_(-1, 1), retain=img(6, 1.0),
             pose=img(5, -1.0), denorm_upper_input=img(3, 1.0), denorm_lower_input=img(3, 1.0),
             denorm_upper_mask=(torch.rand(n, 1, r, r, generator=g) > 0.5).float() * band,
             denorm_lower_mask=(torch.rand(n, 1, r, r, generator=g) > 0.5).float() * band,
             gt_parsing=torch.randint(0, 7, (n, 1, r, r), generator=g).float())
    return {k: v.to(device) for k, v in d.items()}
