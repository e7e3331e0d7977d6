"""Driver for ncu captures of the bounded (shared-memory tile) projection path at 1080p, batch 8.
    python tools/profile_proj.py && ncu --set full --import-source on -k regex:projection_tiled -c 1 -o gpurun_out/prof python tools/profile_proj.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

B, h, w = 8, 1080, 1920
dev = "cuda:0"
inv = synthetic.inv_depth(B, h, w, seed=3).to(dev)
f = synthetic.smooth_flow(B, h, w, 8.0, seed=0).to(dev)
for _ in range(2):
    ops.project_flow(f, inv, 8.0)
torch.cuda.synchronize()
print("ok")
