"""Driver for ncu captures of the frame warp + residual norm at 1080p, batch 8."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

B, h, w = 8, 1080, 1920
f = synthetic.smooth_flow(B, h, w, 8.0, seed=1).cuda()
src = torch.rand((B, h, w, 3), device="cuda") * 255
for _ in range(2):
    ops.warp(src, f, 2, ref=src)
torch.cuda.synchronize()
print("ok")
