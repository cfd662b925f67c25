"""GPU parity of the inference input / output edge (pgpp_u8_to_f32, pgpp_image_to_u8) against the CPU restatement of
test.py:126-147 / :162-166: bit-exact (separately rounded IEEE operations on both sides)."""
import ctypes
import importlib

import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import ref_io

pytestmark = pytest.mark.gpu
load_pkg()
tryon_io = importlib.import_module('pgpp_b200.tryon_io')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'


def _batch(n, h, w, hp, wp, seed, mask_dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    u8 = lambda c, hh, ww: torch.randint(0, 256, (n, c, hh, ww), generator=g, dtype=torch.uint8)
    bit = lambda hh, ww: (torch.rand(n, 1, hh, ww, generator=g) > 0.5)
    return dict(image=u8(3, h, w), pose=u8(3, h, w), norm_img=u8(24, hp, wp), norm_img_lower=u8(21, hp, wp),
                denorm_upper_clothes=u8(3, h, w), denorm_lower_clothes=u8(3, h, w),
                denorm_upper_mask=bit(h, w).to(torch.uint8), denorm_lower_mask=bit(h, w).to(torch.uint8),
                retain_mask=bit(h, w).to(mask_dtype), skin_average=u8(3, h, w), lower_label_map=u8(1, h, w),
                lower_clothes_upper_bound=u8(1, h, w), person_name=['a'] * n)


@pytest.mark.parametrize('shape', [(2, 512, 512, 128, 128), (3, 20, 36, 8, 12), (1, 7, 5, 3, 3), (0, 16, 16, 4, 4)], ids=str)
@pytest.mark.parametrize('mask_dtype', [torch.float32, torch.uint8])
def test_prepare_inputs_is_bit_identical_to_the_reference_expressions(shape, mask_dtype):
    data = _batch(*shape, seed=7, mask_dtype=mask_dtype)
    want = ref_io.prepare_inputs(data)
    got = tryon_io.prepare_inputs(data, DEV)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == torch.float32 and got[k].shape == want[k].shape and torch.equal(got[k].cpu(), want[k]), k


def test_soft_retain_mask():
    data = _batch(2, 64, 48, 8, 8, seed=8)
    data['retain_mask'] = torch.rand(2, 1, 64, 48, generator=torch.Generator().manual_seed(9))
    want = ref_io.prepare_inputs(data)['retain']
    assert torch.equal(tryon_io.prepare_inputs(data, DEV)['retain'].cpu(), want)


@pytest.mark.parametrize('shape', [(4, 3, 512, 512), (2, 3, 17, 9), (1, 3, 1, 1), (0, 3, 8, 8)], ids=str)
def test_images_to_uint8_is_bit_identical(shape):
    g = torch.Generator().manual_seed(11)
    img = torch.randn(*shape, generator=g) * 0.8
    if img.numel() >= 6:
        img.view(-1)[:6] = torch.tensor([-1.0, 1.0, 1.00001, -1.00001, 0.0, 0.999999])
    got = tryon_io.images_to_uint8(img.to(DEV))
    want = ref_io.images_to_uint8(img) if shape[0] else np.zeros((0, shape[2], shape[3], 3), np.uint8)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (shape[0], shape[2], shape[3], 3)
    assert np.array_equal(got.cpu().numpy(), want)
    rgb = tryon_io.images_to_uint8(img.to(DEV), bgr=False)
    assert np.array_equal(rgb.cpu().numpy(), want[..., ::-1])


def test_full_size_round_trip_property():
    """BASELINE batch (32 x 3 x 512 x 512): u8 -> f32 -> u8 is the identity on [1, 254] (the two edges are inverse up to the
    truncation: (x/127.5 - 1 + 1) * 127.5 lands within 1 ulp below x only for values the float grid cannot hit exactly)."""
    g = torch.Generator().manual_seed(12)
    u8 = torch.randint(0, 256, (32, 3, 512, 512), generator=g, dtype=torch.uint8)
    io = custom_ops.get_plugin('io_edge_plugin')
    f = io.u8_to_f32(u8.to(DEV), torch.empty(32, 3, 512, 512, device=DEV))
    back = io.image_to_u8(f, reverse_channels=False).permute(0, 3, 1, 2).cpu()
    want = ((u8.to(torch.float32) / 127.5 - 1 + 1.0) * 127.5).clamp(0, 255).to(torch.uint8)
    assert torch.equal(back, want)
    assert (back.to(torch.int16) - u8.to(torch.int16)).abs().max().item() <= 1


def test_c_abi_rejects_bad_arguments():
    lib = custom_ops.load_library()
    a = torch.zeros(64, dtype=torch.uint8, device=DEV); b = torch.zeros(64, dtype=torch.float32, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.pgpp_u8_to_f32(p(a), 1, 2, 16, p(b), 2, 1, 1, None, None) != 0 and 'channel slice' in lib.pgpp_last_error().decode()
    assert lib.pgpp_u8_to_f32(None, 1, 2, 16, p(b), 2, 0, 1, None, None) != 0
    assert lib.pgpp_image_to_u8(p(b), 1, 17, 2, p(a), 0, None) != 0 and 'bad tensor size' in lib.pgpp_last_error().decode()
    with pytest.raises(RuntimeError):
        custom_ops.get_plugin('io_edge_plugin').u8_to_f32(torch.zeros(1, 1, 4, 4, dtype=torch.uint8), b.view(1, 1, 8, 8))
