"""INTEGRATION.md 2, executed: the REFERENCE's own autograd Functions (`Resample2d`, `ChannelNorm`, `Correlation`:
resample2d.py, channelnorm.py, correlation.py, staged unmodified under oracle/_ref/pyref by oracle/build_ref.py) run
over the drop-in stubs of video_super_resolution_b200/integration/ -- i.e. over libvsr_b200.so -- and are compared with
the same classes running over the reference's own compiled extensions (oracle/_ref/*.so)."""
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref
from video_super_resolution_b200 import integration

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _wrapper(name, backend):
    """The reference's wrapper module `name`.py bound to `backend` in {"b200", "reference"}."""
    ext = name + "_cuda"
    if backend == "b200":
        integration.install()
    else:
        mod = build_ref.load_ref(ext)
        if mod is None:
            pytest.skip(f"oracle/_ref/{ext}.so not built (needs /root/reference in the build container)")
        sys.modules[ext] = mod
    w = build_ref.load_py(name)
    if w is None:
        pytest.skip("oracle/_ref/pyref not staged (needs /root/reference in the build container)")
    return w


def test_reference_resample2d_class_over_the_stub():
    g = torch.Generator().manual_seed(3)
    img = (torch.rand((2, 3, 40, 56), generator=g) * 255).to(DEV)
    flow = ((torch.rand((2, 2, 40, 56), generator=g) - 0.5) * 12).to(DEV)
    gout = torch.randn((2, 3, 40, 56), generator=g).to(DEV)
    res = {}
    for backend in ("b200", "reference"):
        w = _wrapper("resample2d", backend)
        a, f = img.clone().requires_grad_(True), flow.clone().requires_grad_(True)
        out = w.Resample2d()(a, f)                          # resample2d.py:42-51 -> Resample2dFunction.apply
        out.backward(gout)
        res[backend] = (out.detach().cpu().numpy(), a.grad.cpu().numpy(), f.grad.cpu().numpy())
        nearest = w.Resample2d(bilinear=False)(img, flow)
        res[backend] += (nearest.cpu().numpy(),)
    assert np.array_equal(res["b200"][0], res["reference"][0])              # forward: bit for bit
    assert np.array_equal(res["b200"][3], res["reference"][3])              # nearest: bit for bit
    assert np.array_equal(res["b200"][2], res["reference"][2])              # flow gradient: bit for bit
    assert np.abs(res["b200"][1] - res["reference"][1]).max() <= 1e-4       # input gradient: fp32 atomic order


def test_reference_channelnorm_class_over_the_stub():
    x = (torch.randn((2, 3, 33, 47), generator=torch.Generator().manual_seed(1)) * 20).to(DEV)
    gout = torch.randn((2, 1, 33, 47), generator=torch.Generator().manual_seed(2)).to(DEV)
    res = {}
    for backend in ("b200", "reference"):
        w = _wrapper("channelnorm", backend)
        a = x.clone().requires_grad_(True)
        out = w.ChannelNorm()(a)                            # channelnorm.py:32-39
        out.backward(gout)
        res[backend] = (out.detach().cpu().numpy(), a.grad.cpu().numpy())
    assert np.array_equal(res["b200"][0], res["reference"][0])
    assert np.array_equal(res["b200"][1], res["reference"][1])


def test_reference_correlation_class_over_the_stub():
    g = torch.Generator().manual_seed(5)
    a0 = torch.randn((1, 16, 24, 32), generator=g).to(DEV)
    b0 = torch.randn((1, 16, 24, 32), generator=g).to(DEV)
    res = {}
    for backend in ("b200", "reference"):
        w = _wrapper("correlation", backend)
        a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        # FlowNetC's configuration (FlowNetC.py:22)
        out = w.Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)(a, b)
        gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(7)).to(DEV)
        out.backward(gout)
        res[backend] = (out.detach().cpu().numpy(), a.grad.cpu().numpy(), b.grad.cpu().numpy())
    for k in range(3):
        ref = res["reference"][k]
        assert res["b200"][k].shape == ref.shape
        assert np.abs(res["b200"][k] - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())     # fp32 summation order
