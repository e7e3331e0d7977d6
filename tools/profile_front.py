"""Driver for ncu captures of the bandwidth-bound front kernels at 1080p, batch 4 (working set > L2).
    python tools/profile_front.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

B, h, w = 4, 1080, 1920
dev = "cuda:0"
f = synthetic.smooth_flow(B, h, w, 8.0, seed=0).to(dev)
fr = synthetic.random_flow(B, h, w, 64.0, seed=1).to(dev)
inv = synthetic.inv_depth(B, h, w, seed=3).to(dev)
src = (torch.rand((B, h, w, 3)) * 255).to(dev)
lab = synthetic.labels(B, h, w).to(dev)
feat = torch.rand((1, h, w, 32), device=dev)
for _ in range(2):
    ops.project_flow(f)
    ops.project_flow(f, inv)
    ops.project_flow(fr, inv)
    ops.warp(src, f, 2)
    ops.warp(src, f, 2, ref=src)
    ops.warp_labels(lab, f)
    ops.warp(feat, f[:1], 2)
torch.cuda.synchronize()
print("ok")
