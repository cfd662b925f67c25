"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's inference input / output edge.

Follows /root/reference/test.py:126-147 (uint8 -> float32 conversions, retain composition, concatenations) and :162-166
(float image -> uint8 BGR HWC) with the same torch / numpy expressions on CPU tensors.  Only tests/ may import this.
"""
import numpy as np
import torch


def prepare_inputs(data):
    f = lambda k: data[k].to(torch.float32) / 127.5 - 1                       # test.py:126-134,137,140
    image = f('image')
    parts = torch.cat([f('norm_img'), f('norm_img_lower')], dim=1)             # test.py:135
    retain_mask = data['retain_mask'].to(torch.float32)
    retain = image * retain_mask - (1 - retain_mask)                           # test.py:144
    pose = torch.cat([f('pose'), f('lower_label_map'), f('lower_clothes_upper_bound')], dim=1)   # test.py:145
    retain = torch.cat([retain, f('skin_average')], dim=1)                     # test.py:146
    return dict(image=image, parts=parts, pose=pose, retain=retain,
                denorm_upper_clothes=f('denorm_upper_clothes'), denorm_lower_clothes=f('denorm_lower_clothes'),
                denorm_upper_mask=data['denorm_upper_mask'].to(torch.float32),    # test.py:138
                denorm_lower_mask=data['denorm_lower_mask'].to(torch.float32))    # test.py:141


def images_to_uint8(gen_imgs):
    out = []
    for ii in range(gen_imgs.size(0)):                                         # test.py:162-166
        gen_img = gen_imgs[ii].detach().cpu().numpy()
        gen_img = (gen_img.transpose(1, 2, 0) + 1.0) * 127.5
        gen_img = np.clip(gen_img, 0, 255)
        out.append(gen_img.astype(np.uint8)[..., [2, 1, 0]])
    return np.stack(out)
