"""CPU oracle of the full PASTA-GAN++ 512 px generator forward (TEST INFRASTRUCTURE ONLY).

A functional restatement of `GeneratorFull_v20.forward` (training/networks.py:2353-2366) and everything it calls, over a
plain state dict (parameter / buffer names of the reference) and `oracle/ref_ops.py`.  Pinned against the REAL reference:
`oracle/make_golden_generator.py` runs /root/reference's GeneratorFull_v20 (import shims of SURVEY Appendix E, upstream
StyleGAN2-ADA SynthesisLayer injected) with name-seeded weights and stores pooled outputs in tests/golden/generator.npz;
`tests/test_generator_cpu.py` checks this file against them.  Used by tests/, smoke() and bench.py's CPU arm.

Citations are into /root/reference/training/networks.py unless noted.
"""
import math

import torch
import torch.nn.functional as F

from . import ref_ops

SQRT_HALF = math.sqrt(0.5)


def _fc(sd, p, x, activation='linear', lr_multiplier=1.0):
    """FullyConnectedLayer :99-128"""
    w = sd[p + '.weight'] * (lr_multiplier / math.sqrt(sd[p + '.weight'].shape[1]))
    b = sd.get(p + '.bias')
    if b is not None and lr_multiplier != 1:
        b = b * lr_multiplier
    if activation == 'linear' and b is not None:
        return torch.addmm(b.unsqueeze(0), x, w.t())
    return ref_ops.bias_act(x.matmul(w.t()), b, act=activation)


def _conv2d_layer(sd, p, x, f, activation='linear', up=1, down=1, gain=1.0, conv_clamp=None):
    """Conv2dLayer :133-179"""
    w = sd[p + '.weight']
    k = w.shape[-1]
    w = w * (1.0 / math.sqrt(w.shape[1] * k * k))
    y = ref_ops.conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
    act_gain = ref_ops.ACTIVATIONS[activation][1] * gain
    clamp = conv_clamp * gain if conv_clamp is not None else None
    return ref_ops.bias_act(y, sd.get(p + '.bias'), act=activation, gain=act_gain, clamp=clamp)


def _spade_conv(sd, p, x, f, gain=1.0, no_act=False):
    """Spade_Conv2dLayer :1586-1635 (pre-activation relu, bias=False everywhere it is used)"""
    w = sd[p + '.weight']
    k = w.shape[-1]
    w = w * (1.0 / math.sqrt(w.shape[1] * k * k))
    if not no_act:
        x = ref_ops.bias_act(x, sd.get(p + '.bias'), act='relu', gain=ref_ops.ACTIVATIONS['relu'][1] * gain)
    return ref_ops.conv2d_resample(x, w, f=f, padding=k // 2, flip_weight=True)


def _spade_norm(sd, p, x, feats, f):
    """Spade_Norm_Block :1702-1723"""
    normalized = F.instance_norm(x, eps=1e-5)
    actv = torch.relu(_spade_conv(sd, p + '.conv_mlp', feats, f, no_act=True))
    gamma = _spade_conv(sd, p + '.conv_gamma', actv, f, no_act=True)
    beta = _spade_conv(sd, p + '.conv_beta', actv, f, no_act=True)
    return normalized * (1 + gamma) + beta


def _spade_resblock(sd, p, x, feats, f):
    """Spade_ResBlockV4_512 :1859-1904"""
    x = _spade_conv(sd, p + '.conv', x, f, no_act=True)
    y = _spade_conv(sd, p + '.skip', _spade_norm(sd, p + '.spade_skip', x, feats, f), f, gain=SQRT_HALF)
    x = _spade_conv(sd, p + '.conv0', _spade_norm(sd, p + '.spade0', x, feats, f), f)
    x = _spade_conv(sd, p + '.conv1', _spade_norm(sd, p + '.spade1', x, feats, f), f, gain=SQRT_HALF)
    return y + x


def _resblock(sd, p, x, f, activation, down=1):
    """ResBlock :287-316"""
    y = _conv2d_layer(sd, p + '.skip', x, f, down=down, gain=SQRT_HALF)
    x = _conv2d_layer(sd, p + '.conv0', x, f, activation=activation, down=down)
    x = _conv2d_layer(sd, p + '.conv1', x, f, activation=activation, gain=SQRT_HALF)
    return y + x


def _dense(sd, p, x):
    """Dense :391-405: per-pixel Linear, InstanceNorm2d, LeakyReLU(0.01)"""
    out = F.linear(x.permute(0, 2, 3, 1), sd[p + '.linear.weight'], sd[p + '.linear.bias']).permute(0, 3, 1, 2)
    return F.leaky_relu(F.instance_norm(out, eps=1e-5), 0.01)


def const_encoding(sd, pose, f):
    """ConstEncoderNetwork :357-376 (input_nc 5, six down=2 convs)"""
    x = _conv2d_layer(sd, 'const_encoding.model.0', pose, f)
    for i in range(1, 7):
        x = _conv2d_layer(sd, f'const_encoding.model.{i}', x, f, down=2)
    return x


def style_encoding(sd, parts, retain, f):
    """StyleEncoderNetworkV18 :1727-1776"""
    feats = []
    c = retain
    for i in range(4):
        c = _conv2d_layer(sd, f'style_encoding.feat_enc.{i}', c, f, down=(1 if i == 0 else 2))
        feats.append(c)
    x = _conv2d_layer(sd, 'style_encoding.model.0', parts, f)
    idx = 1
    for i in range(6):
        x = _dense(sd, f'style_encoding.model.{idx}', x); idx += 1
        x = _conv2d_layer(sd, f'style_encoding.model.{idx}', x, f, down=(2 if i < 3 else 1)); idx += 1
    x = x.mean(dim=(2, 3))
    return _fc(sd, 'style_encoding.fc', x), feats


def mapping(sd, stylecode, num_ws):
    """MappingNetwork :184-258 with z_dim = 0, num_layers = 1, truncation_psi = 1"""
    y = _fc(sd, 'mapping.embed', stylecode)
    y = y * (y.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
    x = _fc(sd, 'mapping.fc0', y, activation='lrelu', lr_multiplier=0.01)
    return x.unsqueeze(1).repeat(1, num_ws, 1)


def _synthesis_layer(sd, p, x, w, up, f, conv_clamp, noise_mode):
    styles = _fc(sd, p + '.affine', w)
    noise = None
    if noise_mode == 'const' and (p + '.noise_const') in sd:
        noise = sd[p + '.noise_const'] * sd[p + '.noise_strength']
    return ref_ops.synthesis_layer(x, styles, sd[p + '.weight'], sd[p + '.bias'], noise, up, f, conv_clamp=conv_clamp)


def _to_rgb(sd, p, x, w, conv_clamp):
    """ToRGBLayerFull_v1_v4 / _v5 :1910-1967"""
    weight = sd[p + '.weight']
    styles = _fc(sd, p + '.affine', w) * (1.0 / math.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    parsing = None
    if (p + '.m_weight1') in sd:
        parsing = ref_ops.to_rgb(x, styles, sd[p + '.m_weight1'], sd[p + '.m_bias1'], conv_clamp=conv_clamp)
    return ref_ops.to_rgb(x, styles, weight, sd[p + '.bias'], conv_clamp=conv_clamp), parsing


def _block(sd, p, x, img, ws, pose_feat, cat_feats, f, conv_clamp, noise_mode, parsing=None):
    """SynthesisBlockFull_v1_v6.forward :2147-2194 / SynthesisBlockFull_v1_v4.forward :2033-2082 (parsing given)"""
    i = 0
    if (p + '.conv0.weight') in sd:
        x = _synthesis_layer(sd, p + '.conv0', x, ws[:, i], 2, f, conv_clamp, noise_mode); i += 1
        x = _synthesis_layer(sd, p + '.conv1', x, ws[:, i], 1, f, conv_clamp, noise_mode); i += 1
        if x.shape[2] > 32:
            x = _conv2d_layer(sd, p + '.merge_conv', torch.cat([x, cat_feats[str(x.shape[2])]], dim=1), f)
        if parsing is not None:
            x = _spade_resblock(sd, p + '.spade_b512', x, parsing, f)
    else:
        x = _synthesis_layer(sd, p + '.conv1', pose_feat, ws[:, i], 1, f, conv_clamp, noise_mode); i += 1
    if img is not None:
        img = ref_ops.upsample2d(img, f)
    rgb, pred = _to_rgb(sd, p + '.torgb', x, ws[:, i], conv_clamp)
    img = rgb if img is None else img + rgb
    return x, img, pred


def _interp_half(x):
    return F.interpolate(x, scale_factor=0.5)


def _get_spade_feat(sd, mask_512, denorm_mask, denorm_input, f):
    """SynthesisNetworkFull_v18.get_spade_feat :2253-2276"""
    mask_512 = (mask_512 > 0.9).float()
    mask_256 = (_interp_half(mask_512) > 0.9).float()
    denorm_mask_256 = (_interp_half(denorm_mask) > 0.9).float()
    valid = ((mask_256 + denorm_mask_256) == 2.0).float()
    res_mask = mask_256 - valid
    x = denorm_input * mask_512 - (1 - mask_512)
    x = _conv2d_layer(sd, 'synthesis.spade_encoder.0', x, f, activation='relu')
    x = _resblock(sd, 'synthesis.spade_encoder.1', x, f, 'relu')
    feat = _resblock(sd, 'synthesis.spade_encoder.2', x, f, 'relu', down=2)
    valid_sum = (feat * valid).sum(dim=(2, 3), keepdim=True)
    mask_sum = valid.sum(dim=(2, 3), keepdim=True)
    ok = (mask_sum > 10).float()
    mask_sum = mask_sum * ok + (256 * 256) * (1 - ok)
    return feat * (1 - res_mask) + (valid_sum / mask_sum) * res_mask


def synthesis(sd, ws, pose_feat, cat_feats, denorm_upper, denorm_lower, denorm_upper_mask, denorm_lower_mask, gt_parsing,
              conv_clamp=256.0, noise_mode='const'):
    """SynthesisNetworkFull_v18.forward :2279-2327"""
    f = ref_ops.setup_filter([1, 3, 3, 1])
    x = img = pred = None
    w_idx = 0
    res = 8
    while res <= 512:
        p = f'synthesis.b{res}'
        n_conv = 1 if res == 8 else 2
        x, img, pp = _block(sd, p, x, img, ws[:, w_idx:w_idx + n_conv + 1], pose_feat, cat_feats, f, conv_clamp, noise_mode)
        pred = pp if pp is not None else pred
        if res == 256:
            x_256, img_256 = x, img
        if res == 512:
            last_ws = ws[:, w_idx:w_idx + n_conv + 1]
        w_idx += n_conv
        res *= 2
    if gt_parsing is not None:
        parsing_index = gt_parsing
    else:
        parsing_index = torch.argmax(torch.softmax(pred, dim=1), dim=1)[:, None].float()
    upper = ((parsing_index == 1) | (parsing_index == 4)).float()
    lower = ((parsing_index == 2) | (parsing_index == 3)).float()
    up_feat = _get_spade_feat(sd, upper, denorm_upper_mask, denorm_upper, f)
    lo_feat = _get_spade_feat(sd, lower, denorm_lower_mask, denorm_lower, f)
    upper_256 = (_interp_half(upper) > 0.9).float()
    lower_256 = (_interp_half(lower) > 0.9).float()
    spade_feat = up_feat * upper_256 + lo_feat * lower_256
    xs = _spade_resblock(sd, 'synthesis.spade_b256_1', x_256, spade_feat, f)
    xs = _spade_resblock(sd, 'synthesis.spade_b256_2', xs, spade_feat, f)
    _, finetune, _ = _block(sd, 'synthesis.texture_b512', xs, img_256, last_ws, pose_feat, cat_feats, f, conv_clamp, noise_mode,
                            parsing=parsing_index)
    return img, finetune, pred


def generator(sd, c, retain, pose, denorm_upper, denorm_lower, denorm_upper_mask, denorm_lower_mask, gt_parsing=None,
              noise_mode='const'):
    """GeneratorFull_v20.forward :2353-2366 (z has zero width and is not an input here)."""
    f = ref_ops.setup_filter([1, 3, 3, 1])
    pose_feat = const_encoding(sd, pose, f)
    stylecode, feats = style_encoding(sd, c, retain, f)
    ws = mapping(sd, stylecode, num_ws=14)
    cat_feats = {str(t.shape[2]): t for t in feats}
    return synthesis(sd, ws, pose_feat, cat_feats, denorm_upper, denorm_lower, denorm_upper_mask, denorm_lower_mask, gt_parsing,
                     noise_mode=noise_mode)


def synthetic_inputs(n, seed=0):
    """randn.clamp(-1,1) tensors with signal in columns 96..415 (512x320 content), masks rand > 0.5."""
    g = torch.Generator().manual_seed(seed)
    band = torch.zeros(1, 1, 1, 512); band[..., 96:416] = 1

    def img(ch, fill):
        return torch.randn(n, ch, 512, 512, generator=g).clamp_(-1, 1) * band + fill * (1 - band)
    return dict(
        c=torch.randn(n, 45, 128, 128, generator=g).clamp_(-1, 1),
        retain=img(6, 1.0), pose=img(5, -1.0), denorm_upper=img(3, 1.0), denorm_lower=img(3, 1.0),
        denorm_upper_mask=(torch.rand(n, 1, 512, 512, generator=g) > 0.5).float() * band,
        denorm_lower_mask=(torch.rand(n, 1, 512, 512, generator=g) > 0.5).float() * band,
        gt_parsing=torch.randint(0, 7, (n, 1, 512, 512), generator=g).float())


def name_seeded_init(sd_items, scale_bias=0.1):
    """Deterministic weights that do not depend on module construction order: every tensor is drawn from a generator seeded
    by a hash of its name.  `sd_items`: iterable of (name, tensor); tensors are modified in place."""
    import zlib
    for name, t in sd_items:
        if not torch.is_floating_point(t) or name.endswith('resample_filter') or name.endswith('w_avg'):
            continue
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
        v = torch.randn(t.shape, generator=g)
        if name.endswith('noise_strength'):
            v = torch.full(t.shape, 0.1)
        elif name.endswith('affine.bias'):
            v = 1.0 + 0.1 * v
        elif name.endswith('bias') or name.endswith('m_bias1'):
            v = scale_bias * v
        elif name.endswith('linear.weight'):
            v = v / math.sqrt(t.shape[1])
        elif name.startswith('mapping.fc'):
            v = v * 100.0      # lr_multiplier 0.01 parameterisation (FullyConnectedLayer :111)
        with torch.no_grad():
            t.copy_(v)
