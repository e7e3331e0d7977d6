"""Drop-in for the reference's pybind module `channelnorm_cuda` (channelnorm_cuda.cc), over libvsr_b200.so."""
import torch

from ._abi import L, check, stream

_DT = {torch.float32: 0, torch.float16: 1, torch.float64: 2}      # VSR_DTYPE_*: the reference dispatches all three


def forward(input1, output, norm_deg):
    """channelnorm_cuda_forward(input1, output, norm_deg) -- channelnorm.py:14"""
    B, C, H, W = input1.shape
    with __import__("torch").cuda.device(input1.device):
        return check(L().vsr_channelnorm_forward_typed(input1.data_ptr(), output.data_ptr(), B, C, H, W, int(norm_deg),
                                                       _DT[input1.dtype], stream()), "vsr_channelnorm_forward_typed")


def backward(input1, output, grad_output, grad_input1, norm_deg):
    """channelnorm_cuda_backward(input1, output, gradOutput, gradInput1, norm_deg) -- channelnorm.py:26-27"""
    B, C, H, W = input1.shape
    grad_output = grad_output.contiguous()
    with __import__("torch").cuda.device(input1.device):
        return check(L().vsr_channelnorm_backward_typed(input1.data_ptr(), output.data_ptr(), grad_output.data_ptr(),
                                                        grad_input1.data_ptr(), B, C, H, W, int(norm_deg),
                                                        _DT[input1.dtype], stream()), "vsr_channelnorm_backward_typed")
