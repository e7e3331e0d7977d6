"""FlowProjectionModule -- same surface as the reference
(my_packages/FlowProjection/FlowProjectionModule.py:8-33): two HWC frames in, one (h',w',3) fp32
CUDA map out, centre-cropped to multiples of 64 exactly like the reference (:19-25).

What changed is the body.  The reference runs FlowNet2 and colour-codes the flow on the CPU; on
this hot path the flow comes from a pluggable `estimator(input1, input2) -> (h',w',2)` (FlowNet2
itself is out of scope; benchmarks plug synthetic flow) and the module does the
north-star work on the GPU: forward splat with count, normalise, hole fill (ops.project_flow,
SURVEY.md Appendix B).  The three output channels are (projected fx, projected fy, hole mask).

`colour=True` keeps the reference's output instead -- the Middlebury colour code of the estimated flow as an
(h',w',3) fp32 map (:31-32, utils/flow_utils.py:4-24) -- computed on the device (ops.flow_to_image) rather than
through .cpu().numpy() and back.
"""
import torch
from torch.nn.modules.module import Module

from ... import ops
from ...utils.tools import StaticCenterCrop


class FlowProjectionModule(Module):
    def __init__(self, image_size=None, render_size=None, estimator=None, colour=False):
        super(FlowProjectionModule, self).__init__()
        self.colour = colour
        self.cropper = None
        self.image_size = image_size
        self.render_size = render_size
        self.estimator = estimator

    def _crop(self, input1):
        size = tuple(input1.shape[:2])
        render = [(size[0] // 64) * 64, (size[1] // 64) * 64]
        if self.cropper is None or self.image_size != size or self.render_size != render:
            self.image_size, self.render_size = size, render
            self.cropper = StaticCenterCrop(self.image_size, self.render_size)
        return self.cropper

    def project(self, flow):
        """flow (B,h,w,2) or (h,w,2) CUDA fp32 -> dict(proj, wsum, count, hole) from the splat op."""
        squeeze = flow.dim() == 3
        f = flow.unsqueeze(0) if squeeze else flow
        proj, wsum, count, hole = ops.project_flow(f.contiguous())
        if squeeze:
            proj, wsum, count, hole = proj[0], wsum[0], count[0], hole[0]
        return {"proj": proj, "wsum": wsum, "count": count, "hole": hole}

    def colour_code(self, flow, out_size=None):
        """flow (h',w',2) CUDA fp32 -> (h',w',3) fp32 colour image (FlowProjectionModule.py:32), or with
        out_size=(H,W) the (3,H,W) planes after the caller's transpose + nearest resize
        (network/video_super_resolution.py:35)."""
        img, planes = ops.flow_to_image(flow.contiguous(), out_size=out_size, want_u8=out_size is None)
        return img.to(torch.float32) if out_size is None else planes

    def forward(self, input1, input2):
        if self.estimator is None:
            raise RuntimeError("FlowProjectionModule: no flow estimator attached (FlowNet2 is outside the "
                               "B200 hot path); pass estimator=callable or call .project(flow)")
        crop = self._crop(input1)
        flow = self.estimator(crop(input1), crop(input2))          # (h',w',2)
        if self.colour:
            return self.colour_code(flow)
        r = self.project(flow)
        return torch.cat((r["proj"], r["hole"].to(torch.float32).unsqueeze(-1)), dim=-1)
