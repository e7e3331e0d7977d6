"""C2 conv stack (M = 20, 270x480 -> 1080p): one pass timed for every map-chunk size the workspace cap can select.
    gpurun -- 'python tools/time_chunks.py > gpurun_out/time_chunks.log'"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import srfbn_oracle as so  # noqa: E402  (weights initialiser only)
from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule  # noqa: E402

DEV = "cuda:0"
M, h, w = 20, 270, 480
sd = so.init_state_dict(num_maps=M, seed=0, gain=2.3)
x = (torch.rand((M, 3, h, w), generator=torch.Generator().manual_seed(1)) * 255).to(DEV)
full = None
want = None
for chunk in (20, 10, 5, 4, 2, 1):
    cap = None if full is None else int(full * (chunk + 0.5) / M)
    mod = SRProjectionModule(num_maps=M, workspace_cap_bytes=cap)
    mod.load_state_dict(sd)
    u8 = torch.empty((4 * h, 4 * w, 3), dtype=torch.uint8, device=DEV)
    for _ in range(2):
        mod(x, out_u8=u8, want_f32=False)
    ent = next(iter(mod._plans.values()))
    if full is None:
        full = ent["workspace"].numel()
        want = u8.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mod(x, out_u8=u8, want_f32=False)
    e1.record()
    torch.cuda.synchronize()
    print(f"chunk_maps {ent['chunk_maps']:3d}  workspace {ent['workspace'].numel() / 2**30:6.2f} GiB  {e0.elapsed_time(e1) / 5:8.2f} ms per pass"
          f"  identical {bool(torch.equal(u8, want))}", flush=True)
    del mod, ent
    torch.cuda.empty_cache()
