"""Drop-in for the reference's pybind module `correlation_cuda` (correlation_cuda.cc:10-171), over libvsr_b200.so.
The reference's C++ resizes the empty tensors it is handed (`output`, `gradInput1/2`, correlation_cuda.cc:36-42,
:119-125) and zero-fills them; the stub resizes them the same way (the kernels write every element).  The padded
NHWC copies `rInput1/2` are not needed and stay empty."""
import ctypes

from ._abi import L, check, stream


def forward(input1, input2, rbot1, rbot2, output, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
    """correlation_forward_cuda(...) -- correlation.py:26-28"""
    input1, input2 = input1.contiguous(), input2.contiguous()
    B, C, H, W = input1.shape
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(L().vsr_correlation_output_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                           ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)), "vsr_correlation_output_shape")
    output.resize_(B, oc.value, oh.value, ow.value)
    with __import__("torch").cuda.device(input1.device):
        return check(L().vsr_correlation_forward(input1.data_ptr(), input2.data_ptr(), output.data_ptr(), B, C, H, W,
                                                 pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                                                 stream()), "vsr_correlation_forward")


def backward(input1, input2, rbot1, rbot2, grad_output, grad_input1, grad_input2, pad_size, kernel_size, max_displacement,
             stride1, stride2, corr_multiply):
    """correlation_backward_cuda(...) -- correlation.py:42-45"""
    input1, input2, grad_output = input1.contiguous(), input2.contiguous(), grad_output.contiguous()
    B, C, H, W = input1.shape
    grad_input1.resize_(B, C, H, W)
    grad_input2.resize_(B, C, H, W)
    with __import__("torch").cuda.device(input1.device):
        return check(L().vsr_correlation_backward(input1.data_ptr(), input2.data_ptr(), grad_output.data_ptr(),
                                                  grad_input1.data_ptr(), grad_input2.data_ptr(), B, C, H, W, pad_size,
                                                  kernel_size, max_displacement, stride1, stride2, corr_multiply, stream()),
                     "vsr_correlation_backward")
