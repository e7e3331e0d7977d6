"""Small driver for ncu --set full captures of the conv-stack kernels (a full C2 step holds 22 GB of
activations, which ncu would save/restore on every replay).  Same kernels, same tile shapes, fewer
pixels:  python tools/profile_srfbn.py [M h w iters scale]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule  # noqa: E402

M, h, w, iters, scale = (int(a) for a in (sys.argv[1:6] + ["20", "136", "240", "2", "4"][len(sys.argv) - 1:]))
torch.manual_seed(0)
sr = SRProjectionModule(num_maps=M, upscale_factor=scale)
x = (torch.rand((M, 3, h, w)) * 255).cuda()
for _ in range(iters):
    y = sr(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), float(y.mean()))
