"""The caller of the hot path in TRAINING: the PASTA-GAN++ loss of the reference (`training/loss_fullbody.py:32-330`,
class StyleGAN2Loss) restricted to the terms that run without external checkpoints -- adversarial terms for the try-on image,
the refined image and the parsing map, L1, parsing cross-entropy, and the R1 penalty of both discriminators (the
`conv2d_gradfix` double backward).  The VGG / contextual terms need `./checkpoints/vgg19*.pth` (loss_fullbody.py:65,351),
which the reference does not ship; their weights are fixed to 0 here ("training step").

Same constructor arguments, phase names and gradient flow as the reference:
    accumulate_gradients(phase, real_img, gen_z, style_input, retain, pose, denorm_upper_input, denorm_lower_input,
                         denorm_upper_mask, denorm_lower_mask, gt_parsing, sync, gain)
with phase in {Gmain, Greg, Gboth, Dmain, Dreg, Dboth, D_parsingmain, D_parsingreg, D_parsingboth}.  The modules may be
DistributedDataParallel wrappers; `sync=False` suppresses their gradient all-reduce exactly like `misc.ddp_sync`
(torch_utils/misc.py:173-179).  Returns the scalar statistics of the phase as a dict of detached tensors (the reference
reports them through `training_stats`, which is control plane and not rebuilt).
"""
import contextlib

import torch

from ..torch_utils.ops import conv2d_gradfix


@contextlib.contextmanager
def ddp_sync(module, sync):
    if sync or not isinstance(module, torch.nn.parallel.DistributedDataParallel):
        yield
    else:
        with module.no_sync():
            yield


class StyleGAN2Loss:
    def __init__(self, device, G_mapping, G_synthesis, G_const_encoding, G_style_encoding, D, D_parsing, augment_pipe=None,
                 style_mixing_prob=0, r1_gamma=10, pl_weight=0, l1_weight=10, vgg_weight=0, contextual_weight=0, mask_weight=30):
        assert vgg_weight == 0 and contextual_weight == 0, 'the VGG / contextual terms need checkpoints the reference does not ship'
        assert pl_weight == 0, 'path-length regularisation is commented out in the reference (loss_fullbody.py:213-232)'
        self.device = device
        self.G_mapping, self.G_synthesis = G_mapping, G_synthesis
        self.G_const_encoding, self.G_style_encoding = G_const_encoding, G_style_encoding
        self.D, self.D_parsing, self.augment_pipe = D, D_parsing, augment_pipe
        self.style_mixing_prob, self.r1_gamma = style_mixing_prob, r1_gamma
        self.l1_weight, self.mask_weight = l1_weight, mask_weight
        self.ce_parsing = torch.nn.CrossEntropyLoss(ignore_index=255, weight=torch.tensor([1., 3, 4, 4, 4, 4, 4], device=device))

    # ---- loss_fullbody.py:75-112 ----
    def run_G(self, z, c, pose, const_feats, denorm_upper_mask, denorm_lower_mask, denorm_upper_input, denorm_lower_input, gt_parsing, sync):
        cat_feats = {str(f.shape[2]): f for f in const_feats}
        with ddp_sync(self.G_const_encoding, sync):
            pose_feat = self.G_const_encoding(pose)
        with ddp_sync(self.G_mapping, sync):
            ws = self.G_mapping(z, c)
            if self.style_mixing_prob > 0:
                cutoff = torch.empty([], dtype=torch.int64, device=ws.device).random_(1, ws.shape[1])
                cutoff = torch.where(torch.rand([], device=ws.device) < self.style_mixing_prob, cutoff, torch.full_like(cutoff, ws.shape[1]))
                ws[:, cutoff:] = self.G_mapping(torch.randn_like(z), c, skip_w_avg_update=True)[:, cutoff:]
        with ddp_sync(self.G_synthesis, sync):
            img, finetune_img, pred_parsing = self.G_synthesis(ws, pose_feat, cat_feats, denorm_upper_input, denorm_lower_input,
                                                               denorm_upper_mask, denorm_lower_mask, gt_parsing)
        return img, finetune_img, pred_parsing, ws

    def run_D(self, img, pose, c, sync):
        if self.augment_pipe is not None:
            img = self.augment_pipe(img)
        with ddp_sync(self.D, sync):
            return self.D(torch.cat([img, pose[:, 0:3]], dim=1), c)

    def run_D_parsing(self, parsing, pose, c, sync):
        with ddp_sync(self.D_parsing, sync):
            return self.D_parsing(torch.cat([parsing, pose[:, 0:3]], dim=1), c)

    # ---- loss_fullbody.py:115-330 ----
    def accumulate_gradients(self, phase, real_img, gen_z, style_input, retain, pose, denorm_upper_input, denorm_lower_input,
                             denorm_upper_mask, denorm_lower_mask, gt_parsing, sync, gain):
        assert phase in ['Gmain', 'Greg', 'Gboth', 'Dmain', 'Dreg', 'Dboth', 'D_parsingmain', 'D_parsingreg', 'D_parsingboth']
        do_Gmain = phase in ['Gmain', 'Gboth']
        do_Dmain = phase in ['Dmain', 'Dboth']
        do_Dr1 = phase in ['Dreg', 'Dboth'] and self.r1_gamma != 0
        do_DPmain = phase in ['D_parsingmain', 'D_parsingboth']
        do_DPr1 = phase in ['D_parsingreg', 'D_parsingboth'] and self.r1_gamma != 0
        softplus = torch.nn.functional.softplus
        stats = {}

        with ddp_sync(self.G_style_encoding, sync):
            real_c, cat_feats = self.G_style_encoding(style_input, retain)
        gen_c = real_c
        g_args = (gen_z, gen_c, pose, cat_feats, denorm_upper_mask, denorm_lower_mask, denorm_upper_input, denorm_lower_input, gt_parsing)

        if do_Gmain:        # :133-208
            gen_img, gen_finetune_img, pred_parsing, _ = self.run_G(*g_args, sync=sync)
            pred_parsing_onehot = torch.softmax(pred_parsing, dim=1)
            gen_logits = self.run_D(gen_img, pose, gen_c, sync=False)
            gen_finetune_logits = self.run_D(gen_finetune_img, pose, gen_c, sync=False)
            parsing_logits = self.run_D_parsing(pred_parsing_onehot, pose, gen_c, sync=False)
            loss_Gmain = softplus(-gen_logits).mean()
            loss_Gmain_finetune = softplus(-gen_finetune_logits).mean()
            loss_Gmain_parsing = softplus(-parsing_logits).mean()
            loss_L1 = loss_L1_finetune = loss_mask = 0
            if self.l1_weight > 0:
                loss_L1 = torch.nn.functional.l1_loss(gen_img, real_img) * self.l1_weight
                loss_L1_finetune = torch.nn.functional.l1_loss(gen_finetune_img, real_img) * self.l1_weight
            if self.mask_weight > 0:
                loss_mask = torch.mean(self.ce_parsing(pred_parsing, gt_parsing.long()[:, 0])) * self.mask_weight
            loss_G = (loss_Gmain + loss_Gmain_finetune) / 2 + (loss_L1 + loss_L1_finetune) / 2 + loss_mask + loss_Gmain_parsing
            loss_G.mul(gain).backward()
            stats['Loss/G/loss'] = loss_G.detach()

        loss_Dgen_finetune = 0
        if do_Dmain:        # :236-258
            gen_img, gen_finetune_img, _, _ = self.run_G(*g_args, sync=False)
            gen_logits = self.run_D(gen_img, pose, gen_c, sync=False)            # gets synced by the real-image pass
            gen_finetune_logits = self.run_D(gen_finetune_img, pose, gen_c, sync=False)
            loss_Dgen = softplus(gen_logits)
            loss_Dgen_finetune = softplus(gen_finetune_logits)
            ((loss_Dgen.mean() + loss_Dgen_finetune.mean()) / 2).mul(gain).backward()
            stats['Loss/D/gen'] = loss_Dgen_finetune.mean().detach()

        if do_Dmain or do_Dr1:      # :262-284
            real_img_tmp = real_img.detach().requires_grad_(do_Dr1)
            real_logits = self.run_D(real_img_tmp, pose, real_c, sync=sync)
            loss_Dreal = softplus(-real_logits) if do_Dmain else 0
            loss_Dr1 = 0
            if do_Dr1:
                with conv2d_gradfix.no_weight_gradients():
                    r1_grads, = torch.autograd.grad(outputs=[real_logits.sum()], inputs=[real_img_tmp], create_graph=True, only_inputs=True)
                r1_penalty = r1_grads.square().sum([1, 2, 3])
                loss_Dr1 = r1_penalty * (self.r1_gamma / 2)
                stats['Loss/r1_penalty'] = r1_penalty.mean().detach()
            (real_logits * 0 + loss_Dreal + loss_Dr1).mean().mul(gain).backward()
            stats['Loss/scores/real'] = real_logits.mean().detach()

        loss_Dparsing = 0
        if do_DPmain:       # :287-301
            _, _, pred_parsing, _ = self.run_G(*g_args, sync=False)
            parsing_logits = self.run_D_parsing(torch.softmax(pred_parsing, dim=1), pose, gen_c, sync=False)
            loss_Dparsing = softplus(parsing_logits)
            loss_Dparsing.mean().mul(gain).backward()

        if do_DPmain or do_DPr1:    # :305-330
            onehot = torch.cat([(gt_parsing == k).to(gt_parsing.dtype) for k in range(7)], dim=1).detach().requires_grad_(do_DPr1)
            real_parsing_logit = self.run_D_parsing(onehot, pose, real_c, sync=sync)
            loss_DPreal = softplus(-real_parsing_logit) if do_DPmain else 0
            loss_DPr1 = 0
            if do_DPr1:
                with conv2d_gradfix.no_weight_gradients():
                    dp_grads, = torch.autograd.grad(outputs=[real_parsing_logit.sum()], inputs=[onehot], create_graph=True, only_inputs=True)
                dp_penalty = dp_grads.square().sum([1, 2, 3])
                loss_DPr1 = dp_penalty * (self.r1_gamma / 2)
                stats['Loss/DP_r1_penalty'] = dp_penalty.mean().detach()
            (real_parsing_logit * 0 + loss_DPreal + loss_DPr1).mean().mul(gain).backward()
            stats['Loss/scores/real_parsing'] = real_parsing_logit.mean().detach()
        return stats
