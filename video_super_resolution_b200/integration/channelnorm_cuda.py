"""Drop-in for the reference's pybind module `channelnorm_cuda` (channelnorm_cuda.cc), over libvsr_b200.so."""
from ._abi import L, check, stream


def forward(input1, output, norm_deg):
    """channelnorm_cuda_forward(input1, output, norm_deg) -- channelnorm.py:14"""
    B, C, H, W = input1.shape
    with __import__("torch").cuda.device(input1.device):
        return check(L().vsr_channelnorm_forward(input1.data_ptr(), output.data_ptr(), B, C, H, W, int(norm_deg), stream()),
                     "vsr_channelnorm_forward")


def backward(input1, output, grad_output, grad_input1, norm_deg):
    """channelnorm_cuda_backward(input1, output, gradOutput, gradInput1, norm_deg) -- channelnorm.py:26-27"""
    B, C, H, W = input1.shape
    grad_output = grad_output.contiguous()
    with __import__("torch").cuda.device(input1.device):
        return check(L().vsr_channelnorm_backward(input1.data_ptr(), output.data_ptr(), grad_output.data_ptr(),
                                                  grad_input1.data_ptr(), B, C, H, W, int(norm_deg), stream()),
                     "vsr_channelnorm_backward")
