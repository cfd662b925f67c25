"""Mint the generator-level golden fixture from the REAL reference (TEST INFRASTRUCTURE ONLY; build container only).

    python oracle/make_golden_generator.py        # writes tests/golden/generator.npz  (~10 s per forward on 8 threads)

Builds /root/reference's GeneratorFull_v20 with the kwargs of train.py:191-202 through the import shims of SURVEY Appendix E.
The reference tree does not define `SynthesisLayer`; the class below restates the public StyleGAN2-ADA layer from
the behaviour its callers require (Appendix E item 4) using the REFERENCE's own FullyConnectedLayer / modulated_conv2d /
bias_act, and is injected into training.networks before construction.  Weights are name-seeded
(oracle.ref_generator.name_seeded_init) so any implementation with the same parameter names can reproduce them.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from oracle import ref_generator
from oracle.make_golden import OUT, reference_imports


_REF = {}     # reference modules, filled in by main() (module-level class so the reference's persistence layer can pickle it)


class SynthesisLayer(torch.nn.Module):
        def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True, activation='lrelu',
                     resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
            super().__init__()
            self.resolution, self.up, self.use_noise, self.activation, self.conv_clamp = resolution, up, use_noise, activation, conv_clamp
            networks, upfirdn2d, bias_act = _REF['networks'], _REF['upfirdn2d'], _REF['bias_act']
            self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
            self.padding = kernel_size // 2
            self.act_gain = bias_act.activation_funcs[activation].def_gain
            self.affine = networks.FullyConnectedLayer(w_dim, in_channels, bias_init=1)
            self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
            if use_noise:
                self.register_buffer('noise_const', torch.randn([resolution, resolution]))
                self.noise_strength = torch.nn.Parameter(torch.zeros([]))
            self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

        def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1):
            networks, bias_act = _REF['networks'], _REF['bias_act']
            styles = self.affine(w)
            noise = None
            if self.use_noise and noise_mode == 'random':
                noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
            if self.use_noise and noise_mode == 'const':
                noise = self.noise_const * self.noise_strength
            x = networks.modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=noise, up=self.up, padding=self.padding,
                                          resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused_modconv=fused_modconv)
            act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
            return bias_act.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=self.act_gain * gain, clamp=act_clamp)


def pooled(t, k=8):
    return torch.nn.functional.avg_pool2d(t, k).numpy()


def main():
    torch.set_num_threads(8)
    with reference_imports():
        from torch_utils.ops import bias_act, upfirdn2d
        import training.networks as networks
        _REF.update(networks=networks, upfirdn2d=upfirdn2d, bias_act=bias_act)
        networks.SynthesisLayer = SynthesisLayer
        torch.manual_seed(0)
        G = networks.GeneratorFull_v20(z_dim=0, c_dim=512, w_dim=512, img_resolution=512, img_channels=3,
                                       mapping_kwargs=dict(num_layers=1),
                                       synthesis_kwargs=dict(channel_base=32768, channel_max=512, num_fp16_res=3, conv_clamp=256,
                                                             use_noise=True)).eval()
        ref_generator.name_seeded_init(list(G.named_parameters()) + list(G.named_buffers()))
        inp = ref_generator.synthetic_inputs(1, seed=0)
        z = torch.zeros(1, 0)
        out, full = {}, {}
        with torch.no_grad():
            for tag, gt in (('gt', inp['gt_parsing']), ('pred', None)):
                img, fin, pred = G(z, inp['c'], inp['retain'], inp['pose'], inp['denorm_upper'], inp['denorm_lower'],
                                   inp['denorm_upper_mask'], inp['denorm_lower_mask'], gt_parsing=gt, noise_mode='const')
                out[f'{tag}_img_pooled'] = pooled(img); out[f'{tag}_finetune_pooled'] = pooled(fin); out[f'{tag}_parsing_pooled'] = pooled(pred)
                out[f'{tag}_img_crop'] = img[:, :, 200:232, 240:272].numpy(); out[f'{tag}_finetune_crop'] = fin[:, :, 200:232, 240:272].numpy()
                if tag == 'gt':
                    # un-pooled samples over the whole image for the full-resolution parity test: every 3rd (5th) pixel in both
                    # directions, strides coprime with 2 so every polyphase position of the up-sampling layers is hit
                    full['gt_img_s3'] = img[:, :, ::3, ::3].numpy().copy(); full['gt_finetune_s3'] = fin[:, :, ::3, ::3].numpy().copy()
                    full['gt_parsing_s5'] = pred[:, :, ::5, ::5].numpy().copy()
                    full['gt_full_norms'] = np.array([float(img.double().norm()), float(fin.double().norm()), float(pred.double().norm()),
                                                      float(img.abs().max()), float(fin.abs().max()), float(pred.abs().max())])
                out[f'{tag}_stats'] = np.array([float(img.abs().max()), float(img.std()), float(fin.abs().max()), float(fin.std()),
                                                float(pred.abs().max()), float(pred.std())])
        sd = {k: v.detach() for k, v in G.state_dict().items()}
        names = sorted(f'{k}:{tuple(v.shape)}' for k, v in sd.items())
        out['state_dict_names'] = np.array(names)
        out['num_params'] = np.array([sum(p.numel() for p in G.parameters())])
        # cross-check the oracle restatement right here
        with torch.no_grad():
            o_img, o_fin, o_pred = ref_generator.generator(sd, inp['c'], inp['retain'], inp['pose'], inp['denorm_upper'], inp['denorm_lower'],
                                                           inp['denorm_upper_mask'], inp['denorm_lower_mask'], inp['gt_parsing'])
        rel = lambda a, b: float((a - b).norm() / b.norm())
        print('oracle vs reference (gt parsing): img', rel(o_img, img if False else torch.from_numpy(out['gt_img_crop'])) if False else '',
              'pooled img rel', rel(torch.from_numpy(pooled(o_img)), torch.from_numpy(out['gt_img_pooled'])),
              'finetune', rel(torch.from_numpy(pooled(o_fin)), torch.from_numpy(out['gt_finetune_pooled'])),
              'parsing', rel(torch.from_numpy(pooled(o_pred)), torch.from_numpy(out['gt_parsing_pooled'])))
    old = np.load(os.path.join(OUT, 'generator.npz')) if os.path.isfile(os.path.join(OUT, 'generator.npz')) else None
    if old is not None:     # the round-1 fixture stays byte-stable when this run reproduces it
        same = all(np.array_equal(old[k], out[k]) for k in out if k in old)
        print('reproduces the committed generator.npz:', same)
    if old is None or not same:
        np.savez_compressed(os.path.join(OUT, 'generator.npz'), **out)
    np.savez_compressed(os.path.join(OUT, 'generator_fullres.npz'), **full)
    print('params', int(out['num_params'][0]), 'stats', out['gt_stats'])


if __name__ == '__main__':
    main()
