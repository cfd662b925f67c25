"""CPU oracle of the synthesis chain (TEST INFRASTRUCTURE ONLY): the layer sequence of
SynthesisBlockFull_v1_v6.forward (training/networks.py:2147-2194) for blocks b8..b512 plus the texture-branch
block (SynthesisNetworkFull_v18.forward, networks.py:2279-2327, without the SPADE blocks), written as plain
functions over a state dict and `oracle/ref_ops.py`.  Used by tests/, smoke() and bench.py's CPU baseline."""
import math

import torch

from . import ref_ops


def _fc(sd, prefix, x, bias_init_gain=1.0):
    """FullyConnectedLayer with linear activation, lr_multiplier 1 (networks.py:99-128)."""
    w = sd[prefix + '.weight']
    return torch.addmm(sd[prefix + '.bias'].unsqueeze(0), x, (w * (1.0 / math.sqrt(w.shape[1]))).t())


def _synthesis_layer(sd, prefix, x, w, up, f, conv_clamp, noise_mode='const'):
    styles = _fc(sd, prefix + '.affine', w)
    noise = None
    if noise_mode == 'const' and (prefix + '.noise_const') in sd:
        noise = sd[prefix + '.noise_const'] * sd[prefix + '.noise_strength']
    return ref_ops.synthesis_layer(x, styles, sd[prefix + '.weight'], sd[prefix + '.bias'], noise, up, f, conv_clamp=conv_clamp)


def _to_rgb(sd, prefix, x, w, conv_clamp):
    weight = sd[prefix + '.weight']
    styles = _fc(sd, prefix + '.affine', w) * (1.0 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    rgb = ref_ops.to_rgb(x, styles, weight, sd[prefix + '.bias'], conv_clamp=conv_clamp)
    parsing = None
    if (prefix + '.m_weight1') in sd:
        parsing = ref_ops.to_rgb(x, styles, sd[prefix + '.m_weight1'], sd[prefix + '.m_bias1'], conv_clamp=conv_clamp)
    return rgb, parsing


def _merge_conv(sd, prefix, x):
    """Conv2dLayer 1x1, linear (networks.py:133-179)."""
    w = sd[prefix + '.weight']
    y = ref_ops.conv2d_resample(x, w * (1.0 / math.sqrt(w.shape[1] * w.shape[2] ** 2)), padding=0)
    return ref_ops.bias_act(y, sd[prefix + '.bias'])


def _block(sd, name, x, img, ws, pose_feature, cat_feats, f, conv_clamp, noise_mode):
    i = 0
    if (name + '.conv0.weight') in sd:
        x = _synthesis_layer(sd, name + '.conv0', x, ws[:, i], 2, f, conv_clamp, noise_mode); i += 1
        x = _synthesis_layer(sd, name + '.conv1', x, ws[:, i], 1, f, conv_clamp, noise_mode); i += 1
        if (name + '.merge_conv.weight') in sd and cat_feats is not None:
            x = _merge_conv(sd, name + '.merge_conv', torch.cat([x, cat_feats[str(x.shape[2])]], dim=1))
    else:
        x = _synthesis_layer(sd, name + '.conv1', pose_feature, ws[:, i], 1, f, conv_clamp, noise_mode); i += 1
    if img is not None:
        img = ref_ops.upsample2d(img, f)
    rgb, parsing = _to_rgb(sd, name + '.torgb', x, ws[:, i], conv_clamp); i += 1
    img = rgb if img is None else img + rgb
    return x, img, parsing, i


def synthesis_chain(sd, ws, pose_feature, cat_feats=None, img_resolution=512, conv_clamp=256.0, noise_mode='const'):
    """Returns (img, pred_parsing, texture_img) for a state dict of pgpp_b200.training.synthesis.SynthesisChain."""
    f = ref_ops.setup_filter([1, 3, 3, 1])
    x = img = parsing = x_prev = None
    k = 0
    res = 8
    while res <= img_resolution:
        x_prev = x
        x, img, pp, used = _block(sd, f'b{res}', x, img, ws[:, k:], pose_feature, cat_feats, f, conv_clamp, noise_mode)
        parsing = pp if pp is not None else parsing
        k += used
        res *= 2
    _, tex, _, _ = _block(sd, 'texture', x_prev, None, ws[:, k:], None, cat_feats, f, conv_clamp, noise_mode)
    return img, parsing, tex
