"""Bounded vs general vs oracle on one small case; prints where they differ."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from video_super_resolution_b200 import ops, synthetic as syn  # noqa: E402

T, h, w = 7, 70, 200
fl = syn.smooth_flow(T - 1, h, w, 5.0, seed=T + 1)
d = syn.inv_depth(T - 1, h, w, seed=T + 2)
for name, inv in (("depth", d), ("flow", None)):
    want = orc.flow_projection(fl.numpy(), None if inv is None else inv.numpy())
    for md in (None, 5.0):
        got = ops.project_flow(fl.cuda(), None if inv is None else inv.cuda(), md)
        pj = got[0].cpu().numpy()
        diff = np.abs(pj - want[0]).max(axis=-1)
        cnt_ok = np.array_equal(got[2].cpu().numpy(), want[2])
        hole_ok = np.array_equal(got[3].cpu().numpy(), want[3])
        print(name, "max_disp", md, "max|dproj|", diff.max(), "count ok", cnt_ok, "hole ok", hole_ok, "n holes", int(want[3].sum()))
        bad = np.argwhere(diff > 1e-3)
        for b, y, x in bad[:12]:
            print("   ", b, y, x, "hole", want[3][b, y, x], "got", pj[b, y, x], "want", want[0][b, y, x])
