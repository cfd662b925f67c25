"""Input and output edge of the try-on inference loop on the GPU.

`prepare_inputs` is test.py:126-147 - the uint8 -> float32 `/127.5 - 1` conversions, the retain-mask composition and the three
channel concatenations - done by `pgpp_u8_to_f32` launches that write straight into the channel slices of the concatenated
tensors (no `torch.cat` copies, no float32 temporaries).  `images_to_uint8` is test.py:162-166: `(img + 1) * 127.5`, clip,
truncate, RGB -> BGR, HWC.  Both are bit-identical to the reference's expressions (tests/test_gpu_g_io_edge.py).

The field names are the ones the reference's dataset tuple is unpacked into (test.py:121-123).
"""
import torch

from .torch_utils import custom_ops

_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('io_edge_plugin')
    return _plugin


def _dev(t, device):
    return t.to(device, non_blocking=True).contiguous()


def prepare_inputs(data, device):
    """data: dict of the dataset's tensors (uint8 unless noted) -
        image [N,3,H,W], pose [N,Cp,H,W], norm_img [N,Cu,h,w], norm_img_lower [N,Cl,h,w], denorm_upper_clothes [N,3,H,W],
        denorm_lower_clothes [N,3,H,W], denorm_upper_mask [N,1,H,W], denorm_lower_mask [N,1,H,W], retain_mask [N,1,H,W] (float32
        or uint8 0/1), skin_average [N,3,H,W], lower_label_map [N,1,H,W], lower_clothes_upper_bound [N,1,H,W].
    Returns the float32 CUDA tensors test.py feeds the generator: image, parts, pose, retain, denorm_upper_clothes,
    denorm_lower_clothes, denorm_upper_mask, denorm_lower_mask."""
    io = _init()
    d = {k: _dev(v, device) for k, v in data.items() if torch.is_tensor(v)}
    n, _, h, w = d['image'].shape

    def new(c, like):
        return torch.empty([n, c, like.shape[2], like.shape[3]], dtype=torch.float32, device=device)

    out = {}
    out['image'] = io.u8_to_f32(d['image'], new(3, d['image']))                                         # test.py:126
    cu, cl = d['norm_img'].shape[1], d['norm_img_lower'].shape[1]
    parts = new(cu + cl, d['norm_img'])                                                                 # test.py:129-130,135
    io.u8_to_f32(d['norm_img'], parts, 0)
    io.u8_to_f32(d['norm_img_lower'], parts, cu)
    out['parts'] = parts
    cp = d['pose'].shape[1]
    pose = new(cp + d['lower_label_map'].shape[1] + d['lower_clothes_upper_bound'].shape[1], d['pose'])  # test.py:128,133-134,145
    io.u8_to_f32(d['pose'], pose, 0)
    io.u8_to_f32(d['lower_label_map'], pose, cp)
    io.u8_to_f32(d['lower_clothes_upper_bound'], pose, cp + d['lower_label_map'].shape[1])
    out['pose'] = pose
    mask = d['retain_mask']
    mask = mask.to(torch.float32) if mask.dtype != torch.float32 else mask
    retain = new(3 + d['skin_average'].shape[1], d['image'])                                            # test.py:132,143-146
    io.u8_to_f32(d['image'], retain, 0, mask=mask.contiguous())
    io.u8_to_f32(d['skin_average'], retain, 3)
    out['retain'] = retain
    for key in ('denorm_upper_clothes', 'denorm_lower_clothes'):                                        # test.py:137,140
        out[key] = io.u8_to_f32(d[key], new(d[key].shape[1], d[key]))
    for key in ('denorm_upper_mask', 'denorm_lower_mask'):                                              # test.py:138,141
        out[key] = io.u8_to_f32(d[key], new(1, d[key]), normalize=False)
    return out


def images_to_uint8(gen_imgs, bgr=True):
    """float32 [N,3,H,W] generator output -> uint8 [N,H,W,3] device tensor, BGR by default (test.py:162-166)."""
    return _init().image_to_u8(gen_imgs.contiguous(), reverse_channels=bgr)
