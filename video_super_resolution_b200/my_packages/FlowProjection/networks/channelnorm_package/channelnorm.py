"""ChannelNorm -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/channelnorm_package/channelnorm.py:6-39) over the B200 kernel.
x (B,C,H,W) -> (B,1,H,W) = sqrt(sum_c x^2); `norm_deg` is accepted and ignored exactly as the
reference kernel does (channelnorm_kernel.cu:53-59); backward = g*x/(out+1e-9) (:64-96)."""
import torch
from torch import nn

from ..... import ops


class ChannelNormFunction(torch.autograd.Function):
    """apply(input1, norm_deg=2): per-pixel L2 norm over the channel axis, computed by vsr_channelnorm_forward;
    the gradient by vsr_channelnorm_backward.  Nothing is allocated by the caller: the op returns its output."""

    @staticmethod
    def forward(ctx, input1, norm_deg=2):
        if not input1.is_contiguous():
            raise AssertionError("ChannelNorm expects a contiguous (B,C,H,W) tensor, like the reference (:10)")
        norms = ops.channelnorm(input1, norm_deg)
        ctx.norm_deg = norm_deg
        ctx.save_for_backward(input1, norms)
        return norms

    @staticmethod
    def backward(ctx, grad_norms):
        x, norms = ctx.saved_tensors
        grad_x = ops.channelnorm_backward(x, norms, grad_norms.contiguous(), ctx.norm_deg)
        return grad_x, None           # no gradient for norm_deg


class ChannelNorm(nn.Module):
    """Module form used by FlowNet2 (models.py:46-57 of the reference)."""

    def __init__(self, norm_deg=2):
        super().__init__()
        self.norm_deg = norm_deg

    def extra_repr(self):
        return f"norm_deg={self.norm_deg}"

    def forward(self, input1):
        return ChannelNormFunction.apply(input1, self.norm_deg)
