"""ChannelNorm -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/channelnorm_package/channelnorm.py:6-39) over the B200 kernel.
x (B,C,H,W) -> (B,1,H,W) = sqrt(sum_c x^2); `norm_deg` is accepted and ignored exactly as the
reference kernel does (channelnorm_kernel.cu:53-59); backward = g*x/(out+1e-9) (:64-96)."""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import ops


class ChannelNormFunction(Function):
    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        assert input1.is_contiguous()
        ctx.norm_deg = norm_deg
        output = ops.channelnorm(input1, norm_deg)
        ctx.save_for_backward(input1, output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        input1, output = ctx.saved_tensors
        return ops.channelnorm_backward(input1, output, grad_output.contiguous(), ctx.norm_deg), None


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super(ChannelNorm, self).__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)
