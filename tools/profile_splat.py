"""Driver for per-kernel timing of the projection op at 1080p (ncu launch list): smooth, i.i.d. +-64 px, occlusion.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_splat.csv python tools/profile_splat.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

B, h, w = 2, 1080, 1920
dev = "cuda:0"
inv = synthetic.inv_depth(B, h, w, seed=3).to(dev)
for name, f in (("smooth", synthetic.smooth_flow(B, h, w, 8.0, seed=0)), ("iid64", synthetic.random_flow(B, h, w, 64.0, seed=1)),
                ("occl", synthetic.occlusion_scene(B, h, w, shift=64.0, seed=5)[0])):
    f = f.to(dev)
    for _ in range(2):
        ops.project_flow(f, inv)
torch.cuda.synchronize()
print("ok")
