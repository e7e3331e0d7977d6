"""ChannelNorm -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/channelnorm_package/channelnorm.py:6-39) over the B200 kernel.
x (B,C,H,W) -> (B,1,H,W) = sqrt(sum_c x^2); `norm_deg` is accepted and ignored exactly as the
reference kernel does (channelnorm_kernel.cu:53-59).  Forward only (hot path runs under no_grad)."""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import ops


class ChannelNormFunction(Function):
    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        assert input1.is_contiguous()
        ctx.norm_deg = norm_deg
        return ops.channelnorm(input1, norm_deg)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("ChannelNorm backward is outside the B200 hot path (SURVEY.md 8f, rank 2)")


class ChannelNorm(Module):
    def __init__(self, norm_deg=2):
        super(ChannelNorm, self).__init__()
        self.norm_deg = norm_deg

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)
