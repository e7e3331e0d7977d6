"""Resample2d -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/resample2d_package/resample2d.py:6-51) over the B200 kernel.

    Resample2dFunction.apply(input1, input2, kernel_size=1, bilinear=True)
    Resample2d(kernel_size=1, bilinear=True)(input1, input2)

input1 (B,C,H,W), input2 = flow (B,2,H,W), both contiguous fp32 CUDA (the reference asserts
contiguity, :10-11).  The forward arithmetic is the reference kernel's, bit for bit
(csrc/warp.cu); backward is the reference's too (4-tap scatter with int() truncated fractions for
input1, bit-identical flow gradient; resample2d_kernel.cu:75-198).
"""
import torch
from torch import nn

from ..... import ops


def _require_contiguous(t, what):
    if not t.is_contiguous():
        raise AssertionError(f"Resample2d: {what} must be contiguous (the reference asserts the same, resample2d.py:10-11)")


class Resample2dFunction(torch.autograd.Function):
    """Backward warp of `input1` along the pixel-unit flow `input2` (channel 0 horizontal).  The op allocates and
    returns its output (the reference's caller pre-zeroes one and passes it in, :17-19)."""

    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        _require_contiguous(input1, "input1")
        _require_contiguous(input2, "input2 (the flow)")
        ctx.options = (kernel_size, bilinear)
        ctx.save_for_backward(input1, input2)
        return ops.resample2d(input1, input2, kernel_size, bilinear)

    @staticmethod
    def backward(ctx, grad_warped):
        image, flow = ctx.saved_tensors
        kernel_size, bilinear = ctx.options
        grad_image, grad_flow = ops.resample2d_backward(image, flow, grad_warped.contiguous(), kernel_size, bilinear)
        return grad_image, grad_flow, None, None      # kernel_size / bilinear carry no gradient


class Resample2d(nn.Module):
    """Module form used by FlowNet2 (models.py:34-40,86-121 of the reference)."""

    def __init__(self, kernel_size=1, bilinear=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}, bilinear={self.bilinear}"

    def forward(self, input1, input2):
        # the reference makes input1 contiguous here and leaves the flow to the Function's assert (:48-50)
        return Resample2dFunction.apply(input1.contiguous(), input2, self.kernel_size, self.bilinear)
