"""Timeline of CTA 0 of the fused down kernel (knock-out build only: clock64 stamps, see fused_down.cuh VSR_TRACE).
    python -m video_super_resolution_b200.build --knockout
    VSR_B200_LIB=$PWD/video_super_resolution_b200/libvsr_b200_knockout.so VSR_FUSED_TRACE=<nsrc> python tools/fused_trace.py
Prints, for a few steady-state tiles, every stamp relative to the tile's first load, in cycles."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_super_resolution_b200 import _lib  # noqa: E402
from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule  # noqa: E402

NAMES = {0: "load first", 1: "load half0 issued", 2: "load half1 issued", 4: "A0 issued", 5: "A1 issued", 6: "A2 issued",
         7: "A3 issued", 8: "B0 issued", 9: "B1 issued", 10: "B2 issued", 11: "B3 issued", 12: "D_A0 seen", 13: "D_A1 seen",
         14: "D_A2 seen", 15: "D_A3 seen", 16: "conv0 done", 17: "conv1 done", 18: "conv2 done", 19: "conv3 done",
         20: "H free 0", 21: "H free 1", 22: "H free 2", 23: "H free 3", 24: "H3 handed", 25: "D_B seen", 26: "final done"}

M, h, w = 20, 270, 480
torch.manual_seed(0)
sr = SRProjectionModule(num_maps=M)
x = (torch.rand((M, 3, h, w)) * 255).cuda()
for _ in range(2):
    sr(x)
torch.cuda.synchronize()
L = _lib.lib()
buf = np.zeros((1024, 32), dtype=np.int64)
rc = L.vsr_debug_fused_trace(ctypes.c_void_p(buf.ctypes.data))
assert rc == 0, rc
n = int((buf[:, 0] != 0).sum())
print("tiles traced:", n)
per_tile = np.diff(buf[:n, 0])
print("cycles per tile (load first -> next load first): median", int(np.median(per_tile)), "min", int(per_tile.min()), "max", int(per_tile.max()))
for t in (20, 21, 60, 100):
    base = buf[t, 0]
    print(f"--- tile {t} (next tile starts at +{buf[t + 1, 0] - base})")
    for slot, v in sorted(((s, buf[t, s] - base) for s in NAMES if buf[t, s]), key=lambda kv: kv[1]):
        print(f"   {v:8d}  {NAMES[slot]}")
