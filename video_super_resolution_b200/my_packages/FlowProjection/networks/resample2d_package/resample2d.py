"""Resample2d -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/resample2d_package/resample2d.py:6-51) over the B200 kernel.

    Resample2dFunction.apply(input1, input2, kernel_size=1, bilinear=True)
    Resample2d(kernel_size=1, bilinear=True)(input1, input2)

input1 (B,C,H,W), input2 = flow (B,2,H,W), both contiguous fp32 CUDA (the reference asserts
contiguity, :10-11).  The forward arithmetic is the reference kernel's, bit for bit
(csrc/warp.cu).  Every hot-path use is under no_grad (network/video_super_resolution.py:24);
backward is not part of the B200 path and raises.
"""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import ops


class Resample2dFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, bilinear=True):
        assert input1.is_contiguous()
        assert input2.is_contiguous()
        ctx.kernel_size = kernel_size
        ctx.bilinear = bilinear
        return ops.resample2d(input1, input2, kernel_size, bilinear)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("Resample2d backward is outside the B200 hot path (SURVEY.md 8f, rank 2)")


class Resample2d(Module):
    def __init__(self, kernel_size=1, bilinear=True):
        super(Resample2d, self).__init__()
        self.kernel_size = kernel_size
        self.bilinear = bilinear

    def forward(self, input1, input2):
        input1_c = input1.contiguous()
        return Resample2dFunction.apply(input1_c, input2, self.kernel_size, self.bilinear)
