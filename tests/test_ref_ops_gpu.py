"""GPU: the REFERENCE's own CUDA ops (compiled from /root/reference into oracle/_ref by
oracle/build_ref.py) against the C oracle and against the product kernels, bit for bit.
This is the pin of oracle.c::or_resample2d / or_channelnorm (the reference has no tests)."""
import numpy as np
import pytest
import torch

from oracle import build_ref
from oracle import oracle as orc
from video_super_resolution_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(name):
    mod = build_ref.load_ref(name)
    if mod is None:
        pytest.skip(f"oracle/_ref/{name}.so not built (needs /root/reference in the build container)")
    return mod


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (2, 3, 256, 448), (1, 2, 33, 47)])
@pytest.mark.parametrize("bilinear", [True, False])
def test_reference_resample2d_equals_oracle_and_product(shape, bilinear):
    ref = _ref("resample2d_cuda")
    B, C, H, W = shape
    g = torch.Generator().manual_seed(H)
    img = (torch.rand(shape, generator=g) * 255).to(DEV)
    flow = ((torch.rand((B, 2, H, W), generator=g) - 0.5) * 30).to(DEV)
    out = torch.zeros_like(img)                      # resample2d.py:19
    ref.forward(img, flow, out, 1, bilinear)         # resample2d.py:21
    torch.cuda.synchronize()
    want = out.cpu().numpy()
    assert np.array_equal(orc.resample2d_nchw(img.cpu().numpy(), flow.cpu().numpy(), 1, bilinear), want)
    assert np.array_equal(ops.resample2d(img, flow, 1, bilinear).cpu().numpy(), want)


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (2, 2, 100, 37)])
def test_reference_channelnorm_equals_oracle_and_product(shape):
    ref = _ref("channelnorm_cuda")
    x = (torch.randn(shape, generator=torch.Generator().manual_seed(1)) * 20).to(DEV)
    out = torch.zeros((shape[0], 1, shape[2], shape[3]), device=DEV)
    ref.forward(x, out, 2)                           # channelnorm.py:14
    torch.cuda.synchronize()
    want = out.cpu().numpy()
    assert np.array_equal(orc.channelnorm_nchw(x.cpu().numpy()), want)
    assert np.array_equal(ops.channelnorm(x).cpu().numpy(), want)


def test_golden_gpu_fixture_if_present(golden_dir):
    """Fixtures produced by tests/golden/make_golden_gpu.py from the reference ops on a B200."""
    import os
    path = os.path.join(golden_dir, "resample2d_ref.npz")
    if not os.path.exists(path):
        pytest.skip("resample2d_ref.npz not generated yet")
    z = np.load(path)
    for mode, bil in (("bilinear", True), ("nearest", False)):
        got = ops.resample2d(torch.from_numpy(z["img"]).to(DEV), torch.from_numpy(z["flow"]).to(DEV), 1, bil)
        assert np.array_equal(got.cpu().numpy(), z[mode])
    got = ops.channelnorm(torch.from_numpy(z["img"]).to(DEV))
    assert np.array_equal(got.cpu().numpy(), z["channelnorm"])
