"""Layer-by-layer diagnostics of the tcgen05 path (prints error statistics, never asserts):
    gpurun -- 'timeout 300 python tools/debug_srfbn.py > gpurun_out/debug_srfbn.log 2>&1'
Each case runs in its own subprocess with a timeout so that a hung kernel cannot hold the GPU."""
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["fused:1:1:8:16", "fused:2:1:8:16", "fused:3:2:5:7", "fused:6:1:19:33", "deconv:1:19:33:1", "pw:128:32", "pw:1000:64", "pw:4173:192", "deconv:1:8:16:0", "deconv:2:5:7:0", "deconv:1:8:16:1",
         "full:8:16:16:3", "full:20:24:40:3"]


def stats(name, got, want):
    import torch
    got = got.float().cpu()
    want = want.float().cpu()
    err = (got - want).abs()
    fin = torch.isfinite(got).all().item()
    print(f"{name}: finite={fin} max_err={err.max().item():.5g} mean_err={err.mean().item():.5g} "
          f"ref_absmax={want.abs().max().item():.5g} ref_absmean={want.abs().mean().item():.5g} "
          f"got_absmean={got.abs().mean().item():.5g}", flush=True)
    return err


def run_case(case):
    import torch
    import torch.nn.functional as F
    from tests import srfbn_hooks as hk
    dev = "cuda:0"
    parts = case.split(":")
    g = torch.Generator().manual_seed(1)
    if parts[0] == "pw":
        rows, K = int(parts[1]), int(parts[2])
        x = torch.randn((rows, K), generator=g).bfloat16()
        w = (torch.randn((32, K), generator=g) / math.sqrt(K)).bfloat16().float()
        b = torch.randn(32, generator=g) * 0.1
        got = hk.pointwise(x.to(dev), w, b, 0.2)
        want = F.prelu(x.float() @ w.t() + b, torch.tensor([0.2]))
        err = stats(case, got, want)
        print("  per-row-block max err:", [round(err[i:i + 32].max().item(), 3) for i in range(0, min(rows, 256), 32)])
        print("  per-col max err:", [round(v, 3) for v in err.max(0).values.tolist()])
    elif parts[0] == "deconv":
        B, h, w, blk = map(int, parts[1:])
        x = torch.randn((B, h, w, 32), generator=g).bfloat16()
        wt = (torch.randn((32, 32, 8, 8), generator=g) / 16).bfloat16().float()
        b = torch.randn(32, generator=g) * 0.1
        got = hk.deconv(x.to(dev), wt, b, 0.25, block_layout=bool(blk)).cpu()
        want = F.prelu(F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt, b, stride=4, padding=2),
                       torch.tensor([0.25])).permute(0, 2, 3, 1)
        if blk:
            got = hk.from_block(got)
        err = stats(case, got, want)
        print("  err by (y%4,x%4):", [[round(err[:, ry::4, rx::4].max().item(), 3) for rx in range(4)] for ry in range(4)])
    elif parts[0] == "fused":
        nsrc, B, h, w = map(int, parts[1:])
        hr = torch.randn((nsrc, B, 4 * h, 4 * w, 32), generator=g).bfloat16()
        wd = (torch.randn((32, 32, 8, 8), generator=g) / 45).bfloat16().float()
        bd = torch.randn(32, generator=g) * 0.1
        wt = (torch.randn((32, 32 * nsrc), generator=g) / math.sqrt(32 * nsrc)).bfloat16().float()
        bt = torch.randn(32, generator=g) * 0.1
        hrb = torch.stack([hk.to_block(hr[j]) for j in range(nsrc)])
        got = hk.fused_down(hrb.to(dev), wt, bt, 0.3, wd, bd, 0.15)
        x = hr.float().permute(1, 0, 4, 2, 3).reshape(B, nsrc * 32, 4 * h, 4 * w)
        if nsrc > 1:
            x = F.prelu(F.conv2d(x, wt.view(32, 32 * nsrc, 1, 1), bt), torch.tensor([0.3])).bfloat16().float()
        want = F.prelu(F.conv2d(x, wd, bd, stride=4, padding=2), torch.tensor([0.15])).permute(0, 2, 3, 1)
        err = stats(case, got, want)
        print("  err by row:", [round(err[0, y].max().item(), 3) for y in range(min(h, 12))])
        print("  err by col:", [round(err[0, :, xx].max().item(), 3) for xx in range(min(w, 20))])
    elif parts[0] == "full":
        M, h, w, steps = map(int, parts[1:])
        from oracle import srfbn_oracle as so
        from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule
        sd = so.init_state_dict(num_maps=M, seed=3, gain=2.3)
        mod = SRProjectionModule(num_steps=steps, num_maps=M)
        mod.load_state_dict(sd)
        x = torch.rand((M, 3, h, w), generator=g) * 255
        with torch.no_grad():
            wm = so.forward_maps(x, sd, num_steps=steps)
            want = so.fc_fuse(wm, sd)
        gm = mod.premix(x.to(dev)).cpu()
        got = mod(x.to(dev)).cpu()
        e = stats(case + " premix", gm, wm)
        mse = (e ** 2).mean().item()
        print(f"  premix PSNR vs oracle: {10 * math.log10(255 ** 2 / max(mse, 1e-20)):.2f} dB")
        stats(case + " fused", got, want)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for c in CASES:
            try:
                r = subprocess.run([sys.executable, __file__, c], timeout=120, capture_output=True, text=True)
                out = r.stdout + ("\n" + r.stderr[-1500:] if r.returncode else "")
                print(out.strip() + (f"\n  -> exit {r.returncode}" if r.returncode else ""), flush=True)
            except subprocess.TimeoutExpired:
                print(f"{c}: TIMEOUT (kernel hang?)", flush=True)
