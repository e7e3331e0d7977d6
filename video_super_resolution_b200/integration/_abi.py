"""Shared by the three stubs: the library handle, the current stream, error conversion."""
import torch

from .. import _lib


def L():
    return _lib.lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(rc, what):
    if rc:
        raise RuntimeError(f"{what} failed: {_lib.lib().vsr_error_string(rc).decode()} (code {rc})")
    return 1      # what the reference's C++ returns (resample2d_cuda.cc:12,23; correlation_cuda.cc:85,166)
