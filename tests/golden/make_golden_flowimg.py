"""Generates tests/golden/flow2img.npz by RUNNING THE REFERENCE's utils/flow_utils.py:4-24 (flow2img) and the
resize glue of network/video_super_resolution.py:35 (transpose1323 + F.interpolate default nearest) on CPU.

    python tests/golden/make_golden_flowimg.py        (build container only: needs /root/reference)

The arithmetic provider is NumPy; its version is not pinned by the reference.  This fixture pins the behaviour
under the NumPy of this image (2.3.x, NEP 50 promotion): `u / maxrad + np.finfo(float).eps` promotes the fp32
flow to float64, so everything after the normalisation runs in double.  Cases: smooth flow, large i.i.d. flow,
"unknown" (> 1e7) and NaN entries, an all-zero field (0/0 -> NaN -> black).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from utils.flow_utils import flow2img  # noqa: E402
    from utils.tools import transpose1323  # noqa: E402

    rng = np.random.default_rng(11)
    out = {"numpy_version": np.array(np.__version__)}
    cases = {}
    h, w = 48, 64
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    cases["smooth"] = np.stack([3.0 * np.sin(xx / 9.0) + 0.5 * yy / h, 2.0 * np.cos(yy / 7.0) - 1.0], -1).astype(np.float32)
    cases["iid64"] = rng.uniform(-64, 64, (h, w, 2)).astype(np.float32)
    f = rng.normal(0, 2, (h, w, 2)).astype(np.float32)
    f[3, 5, 0] = 2e7
    f[7, 9, 1] = -3e7
    f[11, 2, :] = 1e7          # not > threshold: stays known and dominates maxrad
    cases["unknown"] = f
    f = rng.normal(0, 2, (h, w, 2)).astype(np.float32)
    f[5, 5, 0] = np.nan
    cases["nan"] = f
    cases["zeros"] = np.zeros((h, w, 2), np.float32)
    cases["ragged"] = rng.normal(0, 5, (37, 53, 2)).astype(np.float32)
    for name, flow in cases.items():
        out[f"{name}/flow"] = flow.copy()
        with np.errstate(all="ignore"):
            img = flow2img(flow.copy())                      # the reference mutates its argument
        assert img.dtype == np.uint8
        out[f"{name}/img"] = img
        # video_super_resolution.py:35: interpolate(transpose1323(optical_flow), data_shape), default nearest
        t = torch.tensor(img, dtype=torch.float32)[None]     # FlowProjectionModule.py:32 + torch.stack
        H, W = flow.shape[0] + 14, flow.shape[1] + 32        # crop size -> LR frame size
        out[f"{name}/resized"] = F.interpolate(transpose1323(t), (H, W)).numpy()
    np.savez_compressed(os.path.join(HERE, "flow2img.npz"), **out)
    print("wrote flow2img.npz", {k: v.shape for k, v in out.items() if k.endswith("img")})


if __name__ == "__main__":
    main()
