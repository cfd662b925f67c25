"""`torch.nn.functional.grid_sample` with arbitrarily high order gradients between input and output, on this package's CUDA
kernels.  Drop-in for the reference's torch_utils/ops/grid_sample_gradfix.py (2-D images, mode='bilinear',
padding_mode='zeros', align_corners=False; used by the ADA pipeline, training/augment.py:290-301).

Same public surface: `enabled`, `grid_sample(input, grid)`.  The autograd structure is the reference's
(grid_sample_gradfix.py:41-83): the backward pass is its own Function whose own backward - the gradient of grad_input with
respect to grad_output, needed by the R1 penalty - is again a forward grid_sample.  Where the reference calls
`aten::grid_sampler_2d` / `aten::grid_sampler_2d_backward`, this module calls `pgpp_grid_sample_2d[_backward]`
(csrc/grid_sample.cu).  No fallback: with `enabled = True` a CUDA float32 tensor takes the kernel, anything unsupported raises;
`enabled = False` (or a CPU tensor) is the plain library call, as in the reference.
"""
import torch

from .. import custom_ops

enabled = True      # the reference defaults to False and train.py flips it; here the kernel is the path for CUDA tensors
_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('grid_sample_plugin')
    return True


def grid_sample(input, grid):
    """bilinear / zeros / align_corners=False sampling of `input` [N,C,H,W] at `grid` [N,Ho,Wo,2]"""
    if not _should_use_custom_op(input):
        return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)
    return _Sample.apply(input, grid)


def _should_use_custom_op(input=None):
    return enabled and (input is None or input.device.type == 'cuda')


class _Sample(torch.autograd.Function):
    """out = A(grid) @ input.  Linear in `input`, so its input-gradient is A^T @ grad_out (`_SampleGrad`) and the gradient of THAT
    with respect to grad_out is A again: a forward sample of the incoming second-order gradient."""

    @staticmethod
    def forward(ctx, image, grid):
        assert image.ndim == 4 and grid.ndim == 4
        _init()
        ctx.save_for_backward(image, grid)
        return _plugin.forward(image, grid)

    @staticmethod
    def backward(ctx, d_out):
        image, grid = ctx.saved_tensors
        return _SampleGrad.apply(d_out, image, grid)


class _SampleGrad(torch.autograd.Function):
    """(d_image, d_grid) of `_Sample`; differentiable once more with respect to d_out only (what R1 needs,
    grid_sample_gradfix.py:67-83) - the grid gradient of the second order is not provided, as in the reference."""

    @staticmethod
    def forward(ctx, d_out, image, grid):
        _init()
        ctx.save_for_backward(grid)
        return _plugin.backward(d_out, image, grid)

    @staticmethod
    def backward(ctx, dd_image, dd_grid):
        del dd_grid
        grid, = ctx.saved_tensors
        assert not ctx.needs_input_grad[2]
        dd_out = _Sample.apply(dd_image, grid) if ctx.needs_input_grad[0] else None
        return dd_out, None, None
