"""Small bounded-path projection for compute-sanitizer (racecheck / memcheck):
    compute-sanitizer --tool racecheck python tools/racecheck_proj.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

B, h, w = 2, 200, 300
dev = "cuda:0"
inv = synthetic.inv_depth(B, h, w, seed=3).to(dev)
for f in (synthetic.smooth_flow(B, h, w, 8.0, seed=0), synthetic.random_flow(B, h, w, 8.0, seed=1)):
    ops.project_flow(f.to(dev), inv, 8.0)
    ops.project_flow(f.to(dev), None, 8.0)
torch.cuda.synchronize()
print("ok")
