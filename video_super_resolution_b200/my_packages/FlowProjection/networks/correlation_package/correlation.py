"""Correlation -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/correlation_package/correlation.py:7-64) over the B200 kernel.

    CorrelationFunction.apply(input1, input2, pad_size=3, kernel_size=3, max_displacement=20,
                              stride1=1, stride2=2, corr_multiply=1)
    Correlation(pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1)

FlowNetC uses pad 20, kernel 1, max_displacement 20, strides 1 / 2 (FlowNetC.py:22).  Forward only:
the backward of the cost volume is not built."""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import ops


class CorrelationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2,
                corr_multiply=1):
        return ops.correlation(input1.contiguous(), input2.contiguous(), pad_size, kernel_size, max_displacement,
                               stride1, stride2, corr_multiply)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("Correlation backward is not built (SURVEY.md 8f)")


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super(Correlation, self).__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                         self.stride1, self.stride2, self.corr_multiply)
