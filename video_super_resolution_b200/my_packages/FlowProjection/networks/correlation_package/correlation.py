"""Correlation -- the reference's autograd Function / Module surface
(my_packages/FlowProjection/networks/correlation_package/correlation.py:7-64) over the B200 kernel.

    CorrelationFunction.apply(input1, input2, pad_size=3, kernel_size=3, max_displacement=20,
                              stride1=1, stride2=2, corr_multiply=1)
    Correlation(pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1)

FlowNetC uses pad 20, kernel 1, max_displacement 20, strides 1 / 2 (FlowNetC.py:22).  Backward
(correlation.py:32-47) is built for stride1 = 1, the only case in which the reference's own kernels stay in bounds."""
from torch.autograd import Function
from torch.nn.modules.module import Module

from ..... import ops


class CorrelationFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2,
                corr_multiply=1):
        input1, input2 = input1.contiguous(), input2.contiguous()
        ctx.save_for_backward(input1, input2)
        ctx.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        return ops.correlation(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        g1, g2 = ops.correlation_backward(input1, input2, grad_output.contiguous(), *ctx.cfg)
        return g1, g2, None, None, None, None, None, None


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
