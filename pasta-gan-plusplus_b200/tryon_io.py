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

    class _Edge:
        """uint8 sources take the pgpp_u8_to_f32 kernel; a field the loader delivered as float (lower_label_map = 127.5 for skirts,
        dataset.py:2219) takes the same expression on library ops"""
        @staticmethod
        def u8_to_f32(src, dst, c_off=0, normalize=True, mask=None):
            if src.dtype == torch.uint8:
                return _plugin.u8_to_f32(src, dst, c_off, normalize=normalize, mask=mask)
            assert mask is None
            v = src.to(torch.float32)
            dst[:, c_off:c_off + src.shape[1]] = (v / 127.5 - 1) if normalize else v
            return dst
    io = _Edge

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


FIXTURE_KEYS = {'denorm_upper_img': 'denorm_upper_clothes', 'denorm_lower_img': 'denorm_lower_clothes'}      # dataset.py:2221 names -> test.py:121 names


def load_test_pairs(npz_path, indices=(0,)):
    """Batch of real try-on pairs in the layout the reference's DataLoader hands to test.py:121-123, read from a fixture written by
    the reference's own loader (oracle/make_golden_testpair.py: UvitonDatasetFull_512_test_upper over test_datas/, dataset.py:1952-2223).
    Returns the dict `prepare_inputs` takes (host tensors, uint8 except where the loader itself produced floats)."""
    import numpy as np
    z = np.load(npz_path)
    fields = sorted({k.rsplit('_', 1)[0] for k in z.files if k != 'names'})
    out = {}
    for f in fields:
        arrs = [z[f'{f}_{i}'] for i in indices]
        if any(a.dtype != np.uint8 for a in arrs):
            arrs = [a.astype(np.float32) for a in arrs]
        out[FIXTURE_KEYS.get(f, f)] = torch.from_numpy(np.stack(arrs))
    out['person_name'] = [str(z['names'][i][0]) for i in indices]
    out['clothes_name'] = [str(z['names'][i][1]) for i in indices]
    return out


def tryon(G, inputs, gt_parsing=None, **synthesis_kwargs):
    """test.py:147-160: the generator call of the inference loop on the tensors `prepare_inputs` returns; -> (img, finetune_img,
    pred_parsing).  `z` has zero width (z_dim = 0, test.py:147)."""
    n = inputs['parts'].shape[0]
    z = torch.zeros([n, 0], device=inputs['parts'].device)
    with torch.no_grad():
        return G(z, inputs['parts'], inputs['retain'], inputs['pose'], inputs['denorm_upper_clothes'], inputs['denorm_lower_clothes'],
                 inputs['denorm_upper_mask'], inputs['denorm_lower_mask'], gt_parsing=gt_parsing, **synthesis_kwargs)


def images_to_uint8(gen_imgs, bgr=True):
    """float32 [N,3,H,W] generator output -> uint8 [N,H,W,3] device tensor, BGR by default (test.py:162-166)."""
    return _init().image_to_u8(gen_imgs.contiguous(), reverse_channels=bgr)
