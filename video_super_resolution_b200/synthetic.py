"""Deterministic synthetic inputs of the shapes SURVEY.md 8(d) names (there are no datasets or
pretrained estimators here: flow / depth / masks are synthetic by design of the benchmark).

Everything is generated on the CPU with a seeded torch.Generator (so the CPU oracle and the GPU
see identical bits) and moved by the caller.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gen(seed: int) -> torch.Generator:
    return torch.Generator().manual_seed(int(seed))


def frames(T: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """(T,h,w,3) fp32 in [0,255] -- the reference feeds 0..255 fp32 NHWC (main.py:155-167)."""
    return torch.rand((T, h, w, 3), generator=_gen(seed)) * 255.0


def smooth_flow(B: int, h: int, w: int, max_mag: float = 8.0, seed: int = 0) -> torch.Tensor:
    """(B,h,w,2) fp32: bilinear-upsampled N(0,1) grid scaled to max |.| = max_mag px."""
    gh, gw = max(2, h // 16), max(2, w // 16)
    coarse = torch.randn((B, 2, gh, gw), generator=_gen(seed))
    f = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)
    f = f / f.abs().amax().clamp_min(1e-6) * max_mag
    return f.permute(0, 2, 3, 1).contiguous()


def random_flow(B: int, h: int, w: int, max_mag: float = 64.0, seed: int = 0) -> torch.Tensor:
    """(B,h,w,2) i.i.d. uniform[-max_mag, max_mag] -- the C3 collision stress."""
    return (torch.rand((B, h, w, 2), generator=_gen(seed)) * 2 - 1) * max_mag


def occlusion_scene(B: int, h: int, w: int, shift: float = 64.0, seed: int = 0):
    """C3 'dense occlusion': a foreground rectangle (25% of the pixels, inverse depth 1.0) moves
    +shift px in x over a static background (inverse depth 0.1).  Returns flow (B,h,w,2),
    inv_depth (B,h,w)."""
    flow = torch.zeros((B, h, w, 2))
    inv = torch.full((B, h, w), 0.1)
    g = _gen(seed)
    for b in range(B):
        rh, rw = h // 2, w // 2
        y0 = int(torch.randint(0, max(1, h - rh), (1,), generator=g))
        x0 = int(torch.randint(0, max(1, w - rw), (1,), generator=g))
        flow[b, y0:y0 + rh, x0:x0 + rw, 0] = shift
        inv[b, y0:y0 + rh, x0:x0 + rw] = 1.0
    return flow, inv


def inv_depth(B: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """(B,h,w) uniform (0.05, 1]."""
    return 1.0 - torch.rand((B, h, w), generator=_gen(seed)) * 0.95


def labels(B: int, h: int, w: int, p: float = 0.3, seed: int = 0) -> torch.Tensor:
    """(B,h,w) u8 blobs: thresholded smooth noise, about p of the pixels set."""
    gh, gw = max(2, h // 8), max(2, w // 8)
    coarse = torch.rand((B, 1, gh, gw), generator=_gen(seed))
    s = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)[:, 0]
    thr = torch.quantile(s.flatten()[:: max(1, s.numel() // 65536)], 1.0 - p)
    return (s > thr).to(torch.uint8)


def logits(h: int, w: int, seed: int = 0):
    """Two (h,w) fp32 N(0,2) maps standing in for the OSVOS side outputs."""
    g = _gen(seed)
    return torch.randn((h, w), generator=g) * 2.0, torch.randn((h, w), generator=g) * 2.0
