"""Run ON THE GPU BOX: executes the reference's own CUDA ops (oracle/_ref, compiled from
/root/reference by oracle/build_ref.py) on a small seeded input and writes the outputs to
gpurun_out/resample2d_ref.npz; the file is then committed as tests/golden/resample2d_ref.npz.
    gpurun -- 'python tests/golden/make_golden_gpu.py'
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def main():
    rs = build_ref.load_ref("resample2d_cuda")
    cn = build_ref.load_ref("channelnorm_cuda")
    g = torch.Generator().manual_seed(2024)
    B, C, H, W = 2, 3, 40, 56
    img = torch.rand((B, C, H, W), generator=g) * 255
    flow = (torch.rand((B, 2, H, W), generator=g) - 0.5) * 20
    flow[:, :, 0, 0] = 0.5
    flow[:, :, 1, 1] = -0.5
    flow[:, :, 2, 2] = 300.0
    out = {"img": img.numpy(), "flow": flow.numpy()}
    d_img, d_flow = img.cuda(), flow.cuda()
    for mode, bil in (("bilinear", True), ("nearest", False)):
        o = torch.zeros_like(d_img)
        rs.forward(d_img, d_flow, o, 1, bil)
        out[mode] = o.cpu().numpy()
    o = torch.zeros((B, 1, H, W), device="cuda")
    cn.forward(d_img, o, 2)
    out["channelnorm"] = o.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "resample2d_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
