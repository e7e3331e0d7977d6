"""CPU fp32 restatement of the fusion / upsampling conv stack (SRFBN + per-pixel fc).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header): used by tests/, smoke() and the CPU
baseline legs of bench.py; never imported by the product package.

Follows SRProjectionModule.forward (my_packages/SRProjection/SRProjectionModule.py:133-147),
FeedbackBlock (:7-93) and blocks.py:7-74 of the reference, layer for layer, through the
reference's own arithmetic provider (torch.nn.functional on CPU, fp32), with the INTENDED
dense-concat dataflow of SURVEY.md Appendix C: the reference's FeedbackBlock.forward stages the
concats through `torch.empty` buffers it never fills (:55-59, :70-74), so its whole-network
output depends on uninitialised memory and cannot serve as a golden.  Per-layer behaviour IS
pinned: tests/golden/make_golden.py runs the reference's own ConvBlock/DeconvBlock/MeanShift/fc
modules and this file is checked against those fixtures (tests/test_oracle_golden.py).

Weights are a plain dict in the reference's state-dict naming (Appendix C), so a checkpoint
written by the reference (`main.py:233-237`) loads unchanged.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

RGB_MEAN = (0.4488, 0.4371, 0.4040)  # SRProjectionModule.py:105
# (kernel, stride, padding) of the projection units: x4 is the reference's (SRProjectionModule.py:10-12,
# 101-103); x2 is SRFBN's geometry, used for BASELINE config C4 (the reference has none)
GEOMETRY = {4: (8, 4, 2), 2: (6, 2, 2)}


def init_state_dict(num_maps: int = 8, num_features: int = 32, num_groups: int = 6,
                    seed: int = 0, gain: float = 1.0, upscale: int = 4) -> dict:
    """Random weights with the reference's default initialisers (nn.Conv2d / ConvTranspose2d /
    Linear defaults, PReLU 0.2, MeanShift fixed) and state-dict names.  `gain` scales every conv
    weight: with the default initialisers the signal decays by ~0.6x per layer and the conv branch
    of a 40-layer-deep random network is ~0.03 on a 0..255 image, which would make parity tests
    blind; gain ~2.3 keeps activations O(1..10)."""
    import torch.nn as nn

    g = torch.Generator().manual_seed(seed)
    old = torch.random.get_rng_state()
    torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
    try:
        nf = num_features
        ksp = GEOMETRY[upscale]
        sd = {}

        def put(prefix, mod, act=True):
            sd[prefix + ".0.weight"] = mod.weight.detach().clone() * gain
            sd[prefix + ".0.bias"] = mod.bias.detach().clone()
            if act:
                sd[prefix + ".1.weight"] = torch.full((1,), 0.2)

        mean = torch.tensor(RGB_MEAN)
        sd["sub_mean.weight"] = torch.eye(3).view(3, 3, 1, 1)
        sd["sub_mean.bias"] = -255.0 * mean
        sd["add_mean.weight"] = torch.eye(3).view(3, 3, 1, 1)
        sd["add_mean.bias"] = 255.0 * mean
        put("conv_in", nn.Conv2d(3, 4 * nf, 3, padding=1))
        put("feat_in", nn.Conv2d(4 * nf, nf, 1))
        put("block.compress_in", nn.Conv2d(2 * nf, nf, 1))
        for i in range(num_groups):
            put(f"block.upBlocks.{i}", nn.ConvTranspose2d(nf, nf, *ksp))
            put(f"block.downBlocks.{i}", nn.Conv2d(nf, nf, *ksp))
            if i > 0:
                put(f"block.uptranBlocks.{i - 1}", nn.Conv2d(nf * (i + 1), nf, 1))
                put(f"block.downtranBlocks.{i - 1}", nn.Conv2d(nf * (i + 1), nf, 1))
        put("block.compress_out", nn.Conv2d(num_groups * nf, nf, 1))
        put("out", nn.ConvTranspose2d(nf, nf, *ksp))
        put("conv_out", nn.Conv2d(nf, 3, 3, padding=1), act=False)
        fc0 = nn.Linear(num_maps, 32)
        fc2 = nn.Linear(32, 1)
        sd["fc.0.weight"] = fc0.weight.detach().clone()
        sd["fc.0.bias"] = fc0.bias.detach().clone()
        sd["fc.2.weight"] = fc2.weight.detach().clone()
        sd["fc.2.bias"] = fc2.bias.detach().clone()
        return sd
    finally:
        torch.random.set_rng_state(old)


def _cba(x, sd, prefix, stride=1, padding=0, act=True):
    """ConvBlock: conv(+bias) -> PReLU(1 slope).  ref: blocks.py:7-27,64-74."""
    y = F.conv2d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], stride=stride, padding=padding)
    return F.prelu(y, sd[prefix + ".1.weight"]) if act else y


def _dba(x, sd, prefix, upscale=4):
    """DeconvBlock k8 s4 p2 (x2: k6 s2 p2) -> PReLU.  ref: blocks.py:29-43, SRProjectionModule.py:22-24."""
    _, st, pd = GEOMETRY[upscale]
    y = F.conv_transpose2d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], stride=st, padding=pd)
    return F.prelu(y, sd[prefix + ".1.weight"])


def feedback_block(x, last_hidden, sd, num_groups=6, upscale=4):
    """Intended FeedbackBlock dataflow (SRProjectionModule.py:44-90)."""
    x = _cba(torch.cat((x, last_hidden), 1), sd, "block.compress_in")          # :49-50
    lr = [x]
    hr = []
    for i in range(num_groups):
        ld_l = torch.cat(lr[: i + 1], 1)                                          # :55-59 (intended)
        if i > 0:
            ld_l = _cba(ld_l, sd, f"block.uptranBlocks.{i - 1}")                 # :62-63
        hr.append(_dba(ld_l, sd, f"block.upBlocks.{i}", upscale))                      # :64-65
        ld_h = torch.cat(hr[: i + 1], 1)                                          # :70-74 (intended)
        if i > 0:
            ld_h = _cba(ld_h, sd, f"block.downtranBlocks.{i - 1}")               # :77-78
        lr.append(_cba(ld_h, sd, f"block.downBlocks.{i}", stride=GEOMETRY[upscale][1], padding=2))   # :79-80
    out = _cba(torch.cat(lr[1:], 1), sd, "block.compress_out")                    # :87-88
    return out


def forward_maps(x, sd, num_steps=3, num_groups=6, upscale=4):
    """x (M,3,h,w) fp32 0..255 -> per-map SR output (M,3,4h,4w) before the fc fuse.
    ref: SRProjectionModule.py:134-145 (only the last step's output is kept, :145)."""
    x = F.conv2d(x, sd["sub_mean.weight"], sd["sub_mean.bias"])                   # :135
    inter = F.interpolate(x, scale_factor=upscale, mode="bilinear", align_corners=False)  # :136
    x = _cba(x, sd, "conv_in", padding=1)                                         # :137
    x = _cba(x, sd, "feat_in")                                                    # :138
    hidden = x                                                                    # :45-48 first step
    h = None
    for _ in range(num_steps):
        hidden = feedback_block(x, hidden, sd, num_groups, upscale)                     # :141, :89
        h = hidden
    y = F.conv2d(_dba(h, sd, "out", upscale), sd["conv_out.0.weight"], sd["conv_out.0.bias"], padding=1)
    y = inter + y                                                                 # :142
    y = F.conv2d(y, sd["add_mean.weight"], sd["add_mean.bias"])                   # :143
    return y


def fc_fuse(maps, sd):
    """maps (M,3,H,W) -> (1,3,H,W): Linear(M,32)-ReLU-Linear(32,1)-ReLU over the map axis per
    (channel, pixel).  ref: SRProjectionModule.py:126-131,146; utils/tools.py:118-123."""
    t = maps.permute(1, 2, 3, 0)                                                  # transpose030112
    t = F.relu(F.linear(t, sd["fc.0.weight"], sd["fc.0.bias"]))
    t = F.relu(F.linear(t, sd["fc.2.weight"], sd["fc.2.bias"]))
    return t.permute(3, 0, 1, 2)                                                  # transpose031323


def forward(x, sd, num_steps=3, num_groups=6, upscale=4):
    """Full SRProjectionModule.forward: (M,3,h,w) -> (1,3,4h,4w)."""
    with torch.no_grad():
        return fc_fuse(forward_maps(x, sd, num_steps, num_groups, upscale), sd)


def flops_per_lr_pixel_per_map(num_steps=3, dead_steps_skipped=False):
    """2*MAC count of the stack: 7 347 968 for 3 steps at x4."""
    per_step = 4096 + sum(2048 * (i + 1) * 17 for i in range(1, 6)) + 12 * 131072 + 12288
    tail = 131072 + 27648
    head = 6912 + 8192
    if dead_steps_skipped:
        return head + num_steps * per_step + tail
    return head + num_steps * (per_step + tail)
