"""Device timing of the bandwidth-bound ops at full size (CUDA events, working set > L2).
    gpurun -- 'python tools/time_ops.py > gpurun_out/time_ops.log'
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import ops, synthetic  # noqa: E402

DEV = "cuda:0"
PEAK = 6552.3
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:55s} {ms*1e3:10.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK:6.3f} of measured {PEAK:.0f}", flush=True)


def main():
    h, w = 1080, 1920
    for B in (1, 8):
        P = B * h * w
        flows = {
            "smooth8": synthetic.smooth_flow(B, h, w, 8.0, seed=0),
            "random64": synthetic.random_flow(B, h, w, 64.0, seed=1),
            "occlusion64": synthetic.occlusion_scene(B, h, w, 64.0, seed=2)[0],
        }
        inv = synthetic.inv_depth(B, h, w, seed=3).to(DEV)
        for name, f in flows.items():
            f = f.to(DEV)
            report(f"flow_projection  B={B} {name}", timeit(lambda: ops.project_flow(f)), 21 * P)
            report(f"depth_projection B={B} {name}", timeit(lambda: ops.project_flow(f, inv)), 29 * P)
            if name == "smooth8":
                report(f"flow_projection  B={B} {name} bounded(8)", timeit(lambda: ops.project_flow(f, None, 8.0)), 21 * P)
                report(f"depth_projection B={B} {name} bounded(8)", timeit(lambda: ops.project_flow(f, inv, 8.0)), 29 * P)
                report(f"depth_projection B={B} {name} bounded(16)", timeit(lambda: ops.project_flow(f, inv, 16.0)), 29 * P)
                report(f"depth_projection B={B} {name} bounded(4: broken promise)", timeit(lambda: ops.project_flow(f, inv, 4.0)), 29 * P)
        f = flows["smooth8"].to(DEV)
        src = (torch.rand((B, h, w, 3)) * 255).to(DEV)
        report(f"warp nhwc C=3 exact    B={B}", timeit(lambda: ops.warp(src, f)), 32 * P)
        report(f"warp nhwc C=3 fast     B={B}", timeit(lambda: ops.warp(src, f, 2)), 32 * P)
        report(f"warp nhwc C=3 fast+norm B={B}", timeit(lambda: ops.warp(src, f, 2, ref=src)), 48 * P)
        src_nchw = src.permute(0, 3, 1, 2).contiguous()
        f_nchw = f.permute(0, 3, 1, 2).contiguous()
        report(f"resample2d nchw C=3    B={B}", timeit(lambda: ops.resample2d(src_nchw, f_nchw)), 32 * P)
        lab = synthetic.labels(B, h, w).to(DEV)
        report(f"label warp u8          B={B}", timeit(lambda: ops.warp_labels(lab, f)), 10 * P)
        if B == 1:
            feat = torch.rand((B, h, w, 32), device=DEV)
            report(f"warp nhwc C=32 exact B={B}", timeit(lambda: ops.warp(feat, f)), 264 * P)
            report(f"warp nhwc C=32 fast  B={B}", timeit(lambda: ops.warp(feat, f, 2)), 264 * P)
        # plain copy reference point on this box
        a = torch.empty(P * 8, dtype=torch.float32, device=DEV)
        b = torch.empty_like(a)
        report(f"torch copy {a.numel()*4/1e6:.0f} MB", timeit(lambda: b.copy_(a)), 2 * a.numel() * 4)


if __name__ == "__main__":
    main()
