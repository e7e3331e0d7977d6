"""Device timing of the projection paths at 1080p (CUDA events, batch 8 = working set > L2).
    gpurun -- 'python tools/time_proj.py > gpurun_out/time_proj.log'
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402
from tools.time_ops import timeit, PEAK  # noqa: E402

DEV = "cuda:0"


def report(name, ms, nbytes, B):
    gbs = nbytes / ms / 1e6
    print(f"{name:58s} {ms*1e3/B:8.1f} us/image {gbs:8.1f} GB/s  {gbs/PEAK:6.3f} of measured {PEAK:.0f}  {gbs/8000:6.3f} of 8 TB/s", flush=True)


def main():
    h, w = 1080, 1920
    for B in (8, 1):
        P = B * h * w
        flows = {
            "smooth8": synthetic.smooth_flow(B, h, w, 8.0, seed=0),
            "random8": synthetic.random_flow(B, h, w, 8.0, seed=4),
            "random64": synthetic.random_flow(B, h, w, 64.0, seed=1),
            "occlusion64": synthetic.occlusion_scene(B, h, w, 64.0, seed=2)[0],
        }
        inv = synthetic.inv_depth(B, h, w, seed=3).to(DEV)
        for name, f in flows.items():
            f = f.to(DEV)
            report(f"flow_projection  B={B} {name} general", timeit(lambda: ops.project_flow(f)), 21 * P, B)
            report(f"depth_projection B={B} {name} general", timeit(lambda: ops.project_flow(f, inv)), 29 * P, B)
            if name.endswith("8"):
                report(f"flow_projection  B={B} {name} bounded(8)", timeit(lambda: ops.project_flow(f, None, 8.0)), 21 * P, B)
                report(f"depth_projection B={B} {name} bounded(8)", timeit(lambda: ops.project_flow(f, inv, 8.0)), 29 * P, B)
    # the C2 LR size: 6 pairs of 270 x 480 (launch-latency bound)
    f = synthetic.smooth_flow(6, 270, 480, 8.0, seed=0).to(DEV)
    inv = synthetic.inv_depth(6, 270, 480, seed=3).to(DEV)
    report("depth_projection B=6 270x480 smooth8 bounded(8)", timeit(lambda: ops.project_flow(f, inv, 8.0)), 29 * 6 * 270 * 480, 6)
    report("depth_projection B=6 270x480 smooth8 general", timeit(lambda: ops.project_flow(f, inv)), 29 * 6 * 270 * 480, 6)


if __name__ == "__main__":
    main()
