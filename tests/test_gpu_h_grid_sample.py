"""GPU parity of grid_sample_gradfix (pgpp_grid_sample_2d, pgpp_grid_sample_2d_backward) against the library op the reference
calls (F.grid_sample bilinear / zeros / align_corners=False and its autograd) in float64 on the CPU.  Floating point: the
sample coordinate ((g + 1) * W - 1) / 2 carries W * 2^-24 pixels of fp32 rounding, which the interpolation turns into a value
error of that times the local slope; tolerance 5e-5 (forward) / 2e-4 (gradients) relative to the tensor scale."""
import importlib

import pytest
import torch
import torch.nn.functional as F

from conftest import load_pkg

pytestmark = pytest.mark.gpu
load_pkg()
gs = importlib.import_module('pgpp_b200.torch_utils.ops.grid_sample_gradfix')
custom_ops = importlib.import_module('pgpp_b200.torch_utils.custom_ops')
DEV = 'cuda:0'
TOL = 5e-5


def _close(a, b, tol=TOL):
    scale = max(b.abs().max().item(), 1e-12)
    return (a.detach().cpu().double() - b).abs().max().item() <= tol * scale


def _case(n, c, h, w, ho, wo, seed, spread=1.3):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g)
    grid = (torch.rand(n, ho, wo, 2, generator=g) * 2 - 1) * spread      # beyond [-1, 1]: exercises the zero padding
    return x, grid


@pytest.mark.parametrize('shape', [(2, 3, 64, 64, 64, 64), (1, 5, 17, 9, 23, 31), (3, 1, 1, 1, 4, 4), (2, 4, 8, 8, 1, 1), (0, 3, 8, 8, 4, 4)], ids=str)
def test_forward_and_first_order_gradients(shape):
    x, grid = _case(*shape, seed=21)
    xr, gr = x.double().requires_grad_(True), grid.double().requires_grad_(True)
    yr = F.grid_sample(xr, gr, mode='bilinear', padding_mode='zeros', align_corners=False)
    xg, gg = x.to(DEV).requires_grad_(True), grid.to(DEV).requires_grad_(True)
    yg = gs.grid_sample(xg, gg)
    assert yg.shape == yr.shape and yg.dtype == torch.float32
    if x.numel() == 0:
        return
    assert _close(yg, yr.detach())
    dy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(22))
    gxr, ggr = torch.autograd.grad(yr, [xr, gr], dy.double())
    gxg, ggg = torch.autograd.grad(yg, [xg, gg], dy.to(DEV))
    assert _close(gxg, gxr, 2e-4) and _close(ggg, ggr, 2e-4)


def test_identity_grid_reproduces_the_image_and_borders_are_zero_padded():
    x = torch.randn(2, 3, 16, 24, generator=torch.Generator().manual_seed(23)).to(DEV)
    theta = torch.eye(2, 3).unsqueeze(0).repeat(2, 1, 1)
    grid = F.affine_grid(theta, [2, 3, 16, 24], align_corners=False).to(DEV)
    assert torch.allclose(gs.grid_sample(x, grid), x, atol=1e-4)      # coordinates are integers up to fp32 rounding
    far = torch.full((2, 4, 4, 2), 3.0, device=DEV)
    assert torch.count_nonzero(gs.grid_sample(x, far)) == 0


def test_double_backward_r1_pattern():
    """gradient of |d out / d input|^2 w.r.t. an upstream parameter (the R1 penalty through the augmentation, loss_fullbody.py:264-274)"""
    x, grid = _case(2, 3, 12, 12, 10, 14, seed=24, spread=0.9)

    def run(fn, x, grid, s):
        xi = (x * s).requires_grad_(True)
        y = fn(xi, grid)
        gx, = torch.autograd.grad((y * y).sum(), [xi], create_graph=True)
        return torch.autograd.grad(gx.square().sum(), [s])[0]

    # the library cannot differentiate grid_sampler_2d_backward (that is why the reference has this module); the op is linear in
    # the input, y = s * A x, so loss = |d(y.y)/dx_i|^2 = 4 s^2 |A^T A x|^2 and d loss / d s = 8 s |A^T A x|^2, with A^T from the
    # library's first-order backward in float64
    lib = lambda a, b: F.grid_sample(a, b, mode='bilinear', padding_mode='zeros', align_corners=False)
    x0 = x.double().requires_grad_(True)
    ax = lib(x0, grid.double())
    ata_x, = torch.autograd.grad(ax, [x0], ax.detach())
    ref = 8 * 1.5 * ata_x.square().sum()
    sg = torch.tensor(1.5, device=DEV, requires_grad=True)
    got = run(gs.grid_sample, x.to(DEV), grid.to(DEV), sg)
    assert abs(got.item() - ref.item()) <= 1e-4 * abs(ref.item())


def test_full_size_linearity_and_switches():
    """ADA-sized call (batch 8, 3 x 512 x 512 -> 512 x 512): linear in the input; `enabled = False` routes to the library op"""
    x1, grid = _case(8, 3, 512, 512, 512, 512, seed=25, spread=1.05)
    x2 = torch.randn(x1.shape, generator=torch.Generator().manual_seed(26))
    x1, x2, grid = x1.to(DEV), x2.to(DEV), grid.to(DEV)
    y = gs.grid_sample(0.5 * x1 + x2, grid)
    assert torch.allclose(y, 0.5 * gs.grid_sample(x1, grid) + gs.grid_sample(x2, grid), atol=1e-5)
    launches = custom_ops.launch_count()
    gs.enabled = False
    try:
        lib = gs.grid_sample(x1, grid)
    finally:
        gs.enabled = True
    assert custom_ops.launch_count() == launches
    assert torch.allclose(gs.grid_sample(x1, grid), lib, atol=5e-4)


def test_unsupported_operands_raise():
    x = torch.randn(1, 1, 4, 4, device=DEV)
    with pytest.raises(RuntimeError):
        gs.grid_sample(x.half(), torch.zeros(1, 2, 2, 2, device=DEV, dtype=torch.half))
    with pytest.raises(RuntimeError):
        gs.grid_sample(x, torch.zeros(2, 2, 2, 2, device=DEV))
    assert gs.grid_sample(torch.randn(1, 1, 4, 4), torch.zeros(1, 2, 2, 2)).device.type == 'cpu'     # CPU tensors: library op, as in the reference
