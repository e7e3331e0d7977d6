"""VOSProjectionModule -- same surface as the reference
(my_packages/VOSProjection/VOSProjectionModule.py:6-27): two HWC frames in, an (h,w) {0,1} fp32
mask out.

Body: the pinned elementwise part (sigmoid(a)+sigmoid(b) > 0.7, :22-25) runs on the GPU
(ops.vos_threshold, bit-exact) instead of three host round trips; the OSVOS network that makes
the two side outputs is out of scope, so a pluggable `estimator(input1, input2) -> (a, b)` logits
pair supplies them.  `warp(mask, flow)` is the north star's mask/label warp (nearest, u8,
bit-exact; resample2d_kernel.cu:65-70).  Unlike the reference the mask stays on the device
(the reference returns a CPU tensor only because it thresholds in numpy).
"""
import torch
from torch.nn.modules.module import Module

from ... import ops


class VOSProjectionModule(Module):
    def __init__(self, estimator=None):
        super(VOSProjectionModule, self).__init__()
        self.estimator = estimator

    def threshold(self, logits_a, logits_b):
        return ops.vos_threshold(logits_a.contiguous(), logits_b.contiguous())      # (h,w) u8

    def warp(self, mask_u8, flow):
        """mask (h,w)|(B,h,w) u8, flow (h,w,2)|(B,h,w,2) -> warped labels, same shape."""
        squeeze = mask_u8.dim() == 2
        m = mask_u8.unsqueeze(0) if squeeze else mask_u8
        f = flow.unsqueeze(0) if squeeze else flow
        out = ops.warp_labels(m.contiguous(), f.contiguous())
        return out[0] if squeeze else out

    def forward(self, input1, input2):
        if self.estimator is None:
            raise RuntimeError("VOSProjectionModule: no segmentation estimator attached (OSVOS is outside the "
                               "B200 hot path); pass estimator=callable or call .threshold(a, b)")
        a, b = self.estimator(input1, input2)
        return self.threshold(a, b).to(torch.float32)
