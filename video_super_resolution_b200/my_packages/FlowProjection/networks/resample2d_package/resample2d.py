"""Resample2d -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/resample2d_package/resample2d.py:6-51) over the B200 kernel.

    Resample2dFunction.apply(input1, input2, kernel_size=1, bilinear=True)
    Resample2d(kernel_size=1, bilinear=True)(input1, input2)

input1 (B,C,H,W), input2 = flow (B,2,H,W), both contiguous fp32 CUDA (the reference asserts
contiguity, :10-11).  The forward arithmetic is the reference kernel's, bit for bit
(csrc/warp.cu); backward is the reference's too (4-tap scatter with int() truncated fractions for
input1, bit-identical flow gradient; resample2d_kernel.cu:75-198).
"""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import ops


class Resample2dFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        assert input1.is_contiguous()
        assert input2.is_contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.kernel_size = kernel_size
        ctx.bilinear = bilinear
        return ops.resample2d(input1, input2, kernel_size, bilinear)

    @staticmethod
    def backward(ctx, grad_output):
        grad_output = grad_output.contiguous()
        input1, input2 = ctx.saved_tensors
        g1, g2 = ops.resample2d_backward(input1, input2, grad_output, ctx.kernel_size, ctx.bilinear)
        return g1, g2, None, None


class Resample2d(Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super(Resample2d, self).__init__()
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def forward(self, input1, input2):
        input1_c = input1.contiguous()
        return Resample2dFunction.apply(input1_c, input2, self.kernel_size, self.bilinear)
