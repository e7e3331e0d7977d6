"""DepthProjectionModule -- same surface as the reference
(my_packages/DepthProjection/DepthProjectionModule.py:7-18): `(2,h,w,3)` NHWC frames in, one (h,w)
map out (the caller tiles it x3 with utils.tools.maskprocess, video_super_resolution.py:30-31).

Body: the reference averages two MegaDepth hourglass outputs (:14-17).  Here a pluggable
`estimator(frames) -> (flow (h,w,2), inv_depth (h,w))` supplies the geometry (the hourglass is out
of scope) and the module runs the inverse-depth-weighted splat on the GPU
(ops.project_depth_flow); the returned map is the accumulated inverse depth
of the surfaces landing on each pixel (`wsum`), i.e. the occlusion-resolved depth evidence.
"""
import torch.nn as nn

from ... import ops


class DepthProjectionModule(nn.Module):
    def __init__(self, estimator=None):
        super(DepthProjectionModule, self).__init__()
        self.estimator = estimator

    def project(self, flow, inv_depth):
        """flow (B,h,w,2), inv_depth (B,h,w) -> dict(proj, wsum, count, hole)."""
        squeeze = flow.dim() == 3
        f = flow.unsqueeze(0) if squeeze else flow
        d = inv_depth.unsqueeze(0) if squeeze else inv_depth
        proj, wsum, count, hole = ops.project_depth_flow(f.contiguous(), d.contiguous())
        if squeeze:
            proj, wsum, count, hole = proj[0], wsum[0], count[0], hole[0]
        return {"proj": proj, "wsum": wsum, "count": count, "hole": hole}

    def forward(self, input):
        if self.estimator is None:
            raise RuntimeError("DepthProjectionModule: no depth estimator attached (the MegaDepth hourglass is "
                               "outside the B200 hot path); pass estimator=callable or call .project(flow, inv_depth)")
        flow, inv_depth = self.estimator(input)
        return self.project(flow, inv_depth)["wsum"]
