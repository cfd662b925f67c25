"""Mint the discriminator fixture from the REAL reference (TEST INFRASTRUCTURE ONLY; build container only).

    python oracle/make_golden_discriminator.py        # writes tests/golden/discriminator.npz

/root/reference/training/networks.py's Discriminator (small configurations, name-seeded weights) on the CPU: logits, and the
gradients of the D-step loss with the R1 penalty (loss_fullbody.py:264-274 pattern: softplus(-logits) + gamma/2 * |d logits / d img|^2),
which runs conv2d_gradfix's double backward.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import ref_generator
from oracle.make_golden import OUT, reference_imports

CONFIGS = {
    'resnet_uncond': dict(c_dim=0, img_resolution=32, img_channels=3, channel_base=256, channel_max=16),
    'resnet_cond': dict(c_dim=8, img_resolution=32, img_channels=3, channel_base=256, channel_max=16, conv_clamp=256,
                        mapping_kwargs=dict(num_layers=2)),
    'skip_uncond': dict(c_dim=0, img_resolution=16, img_channels=3, channel_base=128, channel_max=16, architecture='skip',
                        epilogue_kwargs=dict(mbstd_group_size=2)),
}
R1_GAMMA = 10.0


def d_step(D, img, c):
    """(logits, loss, grads of all parameters) of one discriminator step on real images with the R1 penalty"""
    img = img.detach().requires_grad_(True)
    logits = D(img, c)
    r1_grads, = torch.autograd.grad(outputs=[logits.sum()], inputs=[img], create_graph=True, only_inputs=True)
    r1_penalty = r1_grads.square().sum([1, 2, 3])
    loss = (torch.nn.functional.softplus(-logits).squeeze(1) + r1_penalty * (R1_GAMMA / 2)).mean()
    params = [p for p in D.parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, params)
    return logits.detach(), loss.detach(), {n: g for (n, p), g in zip([(n, p) for n, p in D.named_parameters() if p.requires_grad], grads)}


def main():
    torch.set_num_threads(4)
    out = {}
    with reference_imports():
        import training.networks as networks
        for name, cfg in CONFIGS.items():
            D = networks.Discriminator(**cfg).train().requires_grad_(True)
            ref_generator.name_seeded_init(list(D.named_parameters()) + [(n, b) for n, b in D.named_buffers() if 'resample_filter' not in n])
            g = torch.Generator().manual_seed(31)
            img = torch.randn(4, 3, cfg['img_resolution'], cfg['img_resolution'], generator=g).clamp(-1, 1)
            c = torch.randn(4, cfg['c_dim'], generator=g)
            logits, loss, grads = d_step(D, img, c)
            out[f'{name}/img'] = img.numpy(); out[f'{name}/c'] = c.numpy()
            out[f'{name}/logits'] = logits.numpy(); out[f'{name}/loss'] = loss.numpy()
            out[f'{name}/state_names'] = np.array(sorted(f'{k}:{tuple(v.shape)}' for k, v in D.state_dict().items()))
            for k, v in grads.items():
                out[f'{name}/grad/{k}'] = v.numpy()
            print(name, 'logits', logits.flatten().tolist(), 'loss', float(loss), len(grads), 'grads')
    np.savez_compressed(os.path.join(OUT, 'discriminator.npz'), **out)


if __name__ == '__main__':
    main()
