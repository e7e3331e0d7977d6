"""All launches of one general-path depth projection (1080p, batch 2) per field: for `ncu --metrics gpu__time_duration.sum`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

B, h, w = 2, 1080, 1920
dev = "cuda:0"
inv = synthetic.inv_depth(B, h, w, seed=3).to(dev)
for name, f in (("smooth8", synthetic.smooth_flow(B, h, w, 8.0, seed=0)), ("random64", synthetic.random_flow(B, h, w, 64.0, seed=1)),
                ("occlusion64", synthetic.occlusion_scene(B, h, w, 64.0, seed=2)[0])):
    f = f.to(dev)
    ops.project_flow(f, inv)
    torch.cuda.synchronize()
    print(name)
