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
    if _should_use_custom_op(input):
        return _GridSample2dForward.apply(input, grid)
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros', align_corners=False)


def _should_use_custom_op(input=None):
    return enabled and (input is None or input.device.type == 'cuda')


class _GridSample2dForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid):
        assert input.ndim == 4
        assert grid.ndim == 4
        _init()
        output = _plugin.forward(input, grid)
        ctx.save_for_backward(input, grid)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input, grid = ctx.saved_tensors
        grad_input, grad_grid = _GridSample2dBackward.apply(grad_output, input, grid)
        return grad_input, grad_grid


class _GridSample2dBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, grid):
        _init()
        grad_input, grad_grid = _plugin.backward(grad_output, input, grid)
        ctx.save_for_backward(grid)
        return grad_input, grad_grid

    @staticmethod
    def backward(ctx, grad2_grad_input, grad2_grad_grid):
        _ = grad2_grad_grid     # unused, as in the reference (grid_sample_gradfix.py:70)
        grid, = ctx.saved_tensors
        grad2_grad_output = None
        grad2_input = None
        grad2_grid = None
        if ctx.needs_input_grad[0]:
            grad2_grad_output = _GridSample2dForward.apply(grad2_grad_input, grid)
        assert not ctx.needs_input_grad[2]
        return grad2_grad_output, grad2_input, grad2_grid
