"""Correlation -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/correlation_package/correlation.py:7-64) over the B200 kernel.

    CorrelationFunction.apply(input1, input2, pad_size=3, kernel_size=3, max_displacement=20,
                              stride1=1, stride2=2, corr_multiply=1)
    Correlation(pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1)

FlowNetC uses pad 20, kernel 1, max_displacement 20, strides 1 / 2 (FlowNetC.py:22).  Backward
(correlation.py:32-47) is built for stride1 = 1, the only case in which the reference's own kernels stay in bounds."""
import torch
from torch import nn

from ..... import ops

_OPTION_NAMES = ("pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply")


class CorrelationFunction(torch.autograd.Function):
    """Cost volume between two feature maps: (B,C,H,W) x (B,C,H,W) -> (B, D*D, outH, outW).  No padded NHWC
    copies are made (the reference's rbot1 / rbot2, :20-24) and the op allocates its own output."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2,
                corr_multiply=1):
        first, second = input1.contiguous(), input2.contiguous()
        ctx.options = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        ctx.save_for_backward(first, second)
        return ops.correlation(first, second, *ctx.options)

    @staticmethod
    def backward(ctx, grad_volume):
        first, second = ctx.saved_tensors
        grad_first, grad_second = ops.correlation_backward(first, second, grad_volume.contiguous(), *ctx.options)
        return (grad_first, grad_second) + (None,) * len(_OPTION_NAMES)


class Correlation(nn.Module):
    """Module form used by FlowNetC (FlowNetC.py:22,76 of the reference)."""

    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        for name, value in zip(_OPTION_NAMES, (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)):
            setattr(self, name, value)

    def extra_repr(self):
        return ", ".join(f"{n}={getattr(self, n)}" for n in _OPTION_NAMES)

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, *(getattr(self, n) for n in _OPTION_NAMES))
