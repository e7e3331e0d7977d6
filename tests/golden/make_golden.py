"""Generates tests/golden/*.npz by RUNNING THE REFERENCE's own Python code on CPU.

Run once in the build container (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

What is pinned, and how:
  srfbn_layers.npz   every layer of the reference SRProjectionModule (num_features=8 to keep the
                     fixture small; the oracle is generic in num_features), executed through the
                     reference's own nn.Module objects (`blocks.ConvBlock/DeconvBlock/MeanShift`,
                     `fc`, `utils.tools.transpose030112/031323`), one layer at a time, plus its
                     state_dict.  A whole-network reference forward is NOT a usable golden
                     (FeedbackBlock reads torch.empty memory, SRProjectionModule.py:55-59,70-74).
  vos_mask.npz       the numpy lines of VOSProjectionModule.py:22-25 and the MaskedArray fill of
                     network/video_super_resolution.py:58-60 with `utils.tools.maskprocess`,
                     executed verbatim on random logits / images (the module itself cannot be
                     constructed: its pretrained weights are not in the repository).
  layout.npz         utils.tools transpose helpers (Appendix D) on a small index tensor.
The Resample2d / ChannelNorm ops are CUDA-only in the reference; their fixtures are produced on
the GPU box by tests/golden/make_golden_gpu.py from the reference .cu compiled into oracle/_ref.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from my_packages.SRProjection.SRProjectionModule import SRProjectionModule  # noqa: E402
    from utils import tools as rtools  # noqa: E402

    torch.manual_seed(1234)
    nf = 8
    ref = SRProjectionModule(num_features=nf).eval()
    # perturb PReLU slopes / biases so that every parameter matters in the comparison
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.endswith(".1.weight"):
                p.copy_(torch.rand(1) * 0.5 + 0.05)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    out = {"sd/" + k: v.numpy() for k, v in sd.items()}
    g = torch.Generator().manual_seed(7)
    h, w = 5, 6

    def rnd(*shape, scale=1.0):
        return (torch.rand(*shape, generator=g) - 0.5) * 2 * scale

    def rec(name, module, x):
        with torch.no_grad():
            y = module(x)
        out["io/" + name + "/x"] = x.numpy()
        out["io/" + name + "/y"] = y.numpy()

    img = torch.rand(3, 3, h, w, generator=g) * 255.0
    rec("sub_mean", ref.sub_mean, img)
    rec("add_mean", ref.add_mean, rnd(2, 3, 4 * h, 4 * w, scale=100))
    rec("conv_in", ref.conv_in, rnd(2, 3, h, w, scale=100))
    rec("feat_in", ref.feat_in, rnd(2, 4 * nf, h, w))
    rec("block.compress_in", ref.block.compress_in, rnd(2, 2 * nf, h, w))
    for i in range(6):
        rec(f"block.upBlocks.{i}", ref.block.upBlocks[i], rnd(1, nf, h, w))
        rec(f"block.downBlocks.{i}", ref.block.downBlocks[i], rnd(1, nf, 4 * h, 4 * w))
        if i > 0:
            rec(f"block.uptranBlocks.{i - 1}", ref.block.uptranBlocks[i - 1], rnd(1, nf * (i + 1), h, w))
            rec(f"block.downtranBlocks.{i - 1}", ref.block.downtranBlocks[i - 1], rnd(1, nf * (i + 1), 2 * h, 2 * w))
    rec("block.compress_out", ref.block.compress_out, rnd(2, 6 * nf, h, w))
    rec("out", ref.out, rnd(2, nf, h, w))
    rec("conv_out", ref.conv_out, rnd(2, nf, 4 * h, 4 * w))
    # the bilinear skip exactly as called at SRProjectionModule.py:136
    xs = rnd(2, 3, h, w, scale=100)
    out["io/skip/x"] = xs.numpy()
    out["io/skip/y"] = torch.nn.functional.interpolate(
        xs, scale_factor=ref.upscale_factor, mode="bilinear", align_corners=False).numpy()
    # the fc fuse exactly as written at SRProjectionModule.py:146 (8 maps)
    maps = rnd(8, 3, 7, 9, scale=100)
    with torch.no_grad():
        fused = rtools.transpose031323(ref.fc(rtools.transpose030112(maps))).squeeze()
    out["io/fc/x"] = maps.numpy()
    out["io/fc/y"] = fused.numpy()
    np.savez_compressed(os.path.join(HERE, "srfbn_layers.npz"), **out)

    # ---- VOS mask lines -------------------------------------------------------------------
    rng = np.random.default_rng(11)
    la = rng.normal(0, 2, (2, 1, 33, 47)).astype(np.float32)           # outputs[-1] (2,1,h,w)
    preds = np.transpose(la, (0, 2, 3, 1))                             # VOSProjectionModule.py:21
    preds = [np.squeeze(1 / (1 + np.exp(-pred))) for pred in preds]     # :22
    pred = preds[0] + preds[1]                                          # :23
    pred[pred > 0.7] = 1                                                # :24
    pred[pred <= 0.7] = 0                                               # :25
    mask_t = torch.tensor(pred)                                         # :26
    vosmask = rtools.maskprocess(mask_t)                                # video_super_resolution.py:54
    image = (rng.random((3, 33, 47)) * 255).astype(np.float32)
    filled = np.ma.MaskedArray(image, vosmask, fill_value=0).filled()  # :58-59
    np.savez_compressed(os.path.join(HERE, "vos_mask.npz"), logits=la, mask=pred.astype(np.float32),
                        image=image, filled=filled.astype(np.float32))

    # ---- layout helpers -------------------------------------------------------------------
    t4 = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).view(2, 3, 4, 5)
    t3 = torch.arange(3 * 4 * 5, dtype=torch.float32).view(3, 4, 5)
    lay = {"t4": t4.numpy(), "t3": t3.numpy()}
    for name in ["transpose1323", "transpose1223", "transpose1312", "transpose030112", "transpose031323"]:
        lay[name] = getattr(rtools, name)(t4).contiguous().numpy()
    lay["transpose1201"] = rtools.transpose1201(t3).contiguous().numpy()
    lay["maskprocess"] = rtools.maskprocess(t3[0]).numpy()
    np.savez_compressed(os.path.join(HERE, "layout.npz"), **lay)
    for f in ["srfbn_layers.npz", "vos_mask.npz", "layout.npz"]:
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
