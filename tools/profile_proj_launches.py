"""All launches of one bounded depth projection (1080p, batch 8): for `ncu --metrics gpu__time_duration.sum`, and a
plain CUDA-event timing of the call with and without the stages after the tile kernel (env VSR_PROJ_ONLY_TILE)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402
from tools.time_ops import timeit  # noqa: E402

B, h, w = 8, 1080, 1920
dev = "cuda:0"
inv = synthetic.inv_depth(B, h, w, seed=3).to(dev)
f = synthetic.smooth_flow(B, h, w, 8.0, seed=0).to(dev)
for _ in range(3):
    ops.project_flow(f, inv, 8.0)
torch.cuda.synchronize()
print("per call ms", timeit(lambda: ops.project_flow(f, inv, 8.0)))
