#!/usr/bin/env python
"""Timings of BASELINE.json's other configurations (bench.py keeps the contract line on C2):

    python tools/bench_configs.py --config c3            # DepthProjection splat, 1080p, +-64 px, GB/s vs HBM
    python tools/bench_configs.py --config c4            # 2x VSR 1080p -> 4K, T=5 (M=14), one window per GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        tools/bench_configs.py --config c5 [--frames 300] [--exact-chunks]
                                                         # 300-frame 720x360 -> 2880x1440 sequence, chunk-sharded

One JSON line per run on stdout (rank 0).  Device timing with CUDA events, barrier + synchronize on both
sides, max over ranks; synthetic seeded inputs."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"]))
    except Exception:
        return 6650.0, 1400.0


def setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("VSR_NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # the NCCL banner must not land on stdout
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, dev


def barrier(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world, dev):
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gained(sr, gain=2.3):
    import torch
    with torch.no_grad():
        for name, p in sr.named_parameters():
            if name.endswith(".0.weight") and not name.startswith(("sub_mean", "add_mean")):
                p.mul_(gain)
    return sr


def run_c3(args):
    """DepthProjection splat only, 1080p, B=8 images rotated so the inputs (8 x 25 MB) + outputs exceed L2."""
    import torch
    from video_super_resolution_b200 import ops, synthetic as syn
    rank, world, dev = setup()
    hbm, _ = peaks()
    h, w, B = 1080, 1920, 8
    out = {}
    cases = {"iid_pm64": lambda: (syn.random_flow(B, h, w, 64.0, seed=8), syn.inv_depth(B, h, w, seed=9)),
             "dense_occlusion_64": lambda: syn.occlusion_scene(B, h, w, shift=64.0, seed=5),
             "smooth_pm8": lambda: (syn.smooth_flow(B, h, w, 8.0, seed=1), syn.inv_depth(B, h, w, seed=2))}
    for name, make in cases.items():
        flow, inv = make()
        flow, inv = flow.to(dev), inv.to(dev)
        for _ in range(5):
            ops.project_depth_flow(flow, inv)
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ops.project_depth_flow(flow, inv)
        e1.record()
        barrier(world)
        ms = e0.elapsed_time(e1) / args.steps
        gb = 29.0 * B * h * w / ms / 1e6
        out[name] = {"us_per_image": ms * 1e3 / B, "GBps": gb, "frac_hbm_measured": gb / hbm, "frac_hbm_8TBs": gb / 8000.0}
    if rank == 0:
        print(json.dumps({"config": "C3: DepthProjection splat + normalise + hole fill, 1920x1080, batch 8 (29 B/pixel)",
                          "metric": "depth_projection_GBps", "unit": "GB/s", "n_gpus": world, "steps": args.steps,
                          "hbm_peak_measured": hbm, "cases": out}), flush=True)


def run_c4(args):
    """2x VSR 1920x1080 -> 3840x2160, 5-frame window (M=14), one window per GPU (weak scaling)."""
    import torch
    from video_super_resolution_b200 import synthetic as syn
    from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule
    from video_super_resolution_b200.pipeline import WarpFusePipeline
    rank, world, dev = setup()
    _, tf = peaks()
    T, h, w, s = 5, 1080, 1920, 2
    M = 3 * T - 1
    torch.manual_seed(0)
    sr = gained(SRProjectionModule(num_maps=M, upscale_factor=s))
    pipe = WarpFusePipeline(T, h, w, sr, s, device=dev)
    la, lb = syn.logits(h, w, seed=3 + rank)
    inp = [syn.frames(T, h, w, seed=rank).to(dev), syn.smooth_flow(T - 1, h, w, 8.0, seed=1 + rank).to(dev),
           syn.inv_depth(T - 1, h, w, seed=2 + rank).to(dev), la.to(dev), lb.to(dev)]
    for _ in range(max(args.warmup, 1)):
        pipe.step(*inp)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.step(*inp)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, world, dev)
    sr.profile(True)
    sr(pipe.stack)
    prof = sr.profile_read()
    sr.profile(False)
    if rank == 0:
        flops = 2 * M * h * w * 3.575e6     # 3.575 MFLOP per LR pixel per map at x2, two passes
        print(json.dumps({"config": "C4: 2x VSR 1920x1080->3840x2160, 5-frame window (M=14), one window per GPU; "
                                    "SRFBN k6 s2 p2 geometry, layered tcgen05 kernels",
                          "metric": "sr_frames_per_s_2x_4k_out", "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                          "value": world / (ms / 1e3), "ms_per_step": ms, "scaling": "weak",
                          "tflops_per_step": flops / 1e12, "step_tensor_frac": flops / ms / 1e9 / tf,
                          "kernels_one_pass": {k: {"ms": round(v["ms"], 2), "launches": v["launches"],
                                                   "GBps": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                                                   "TFLOPs": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)}
                                               for k, v in prof.items() if v["launches"]}}), flush=True)


def run_c2t3(args):
    """C2's frame size with the REFERENCE's window: 480x270 -> 1920x1080, T = 3 (M = 8 maps, SRProjectionModule.py:127)."""
    import torch
    from video_super_resolution_b200 import synthetic as syn
    from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule
    from video_super_resolution_b200.pipeline import WarpFusePipeline
    rank, world, dev = setup()
    _, tf = peaks()
    T, h, w, s = 3, 270, 480, 4
    M = 3 * T - 1
    torch.manual_seed(0)
    sr = gained(SRProjectionModule(num_maps=M))
    pipe = WarpFusePipeline(T, h, w, sr, s, device=dev)
    la, lb = syn.logits(h, w, seed=3 + rank)
    inp = [syn.frames(T, h, w, seed=rank).to(dev), syn.smooth_flow(T - 1, h, w, 8.0, seed=1 + rank).to(dev),
           syn.inv_depth(T - 1, h, w, seed=2 + rank).to(dev), la.to(dev), lb.to(dev)]
    for _ in range(max(args.warmup, 3)):
        pipe.step(*inp)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.step(*inp)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, world, dev)
    if rank == 0:
        from oracle import srfbn_oracle as so   # FLOP accounting only (the out / conv_out of the two dead feedback steps are skipped)
        flops = 2 * M * h * w * so.flops_per_lr_pixel_per_map(dead_steps_skipped=True)
        print(json.dumps({"config": "C2 frame size with the reference's 3-frame window (M=8): 4x VSR 480x270->1920x1080, "
                                    "one window per GPU", "metric": "sr_frames_per_s_4x_1080p_out_T3", "unit": "frames/s",
                          "n_gpus": world, "steps": args.steps, "value": world / (ms / 1e3), "ms_per_step": ms,
                          "scaling": "weak", "tflops_per_step": flops / 1e12,
                          "step_tensor_frac": flops / ms / 1e9 / tf}), flush=True)


def run_c5(args):
    """300-frame 720x360 -> 2880x1440 (4x) sequence, T=3 windows (M=8), the reference's 20 chunks with the
    recurrence reset at chunk starts (main.py:196), sharded over the ranks; one NCCL gather of the u8 frames."""
    import torch
    from video_super_resolution_b200 import synthetic as syn
    from video_super_resolution_b200.network.video_super_resolution import VSR
    from video_super_resolution_b200.pipeline import gather_frames_ragged, run_sequence, shard_chunks, shard_chunks_even
    from video_super_resolution_b200.utils.video_utils import chunk_windows
    rank, world, dev = setup()
    T, h, w, s, n = 3, 360, 720, 4, args.frames
    torch.manual_seed(0)
    vsr = VSR(window=T)
    gained(vsr.model)
    chunks = chunk_windows(n, T)
    mine = (shard_chunks if args.exact_chunks else shard_chunks_even)(chunks, world, rank)
    lo = min(c.start for c in mine)
    hi = max(c.stop for c in mine) + T - 1                # frames this rank needs (T-1 halo frames, replicated)
    # host side: the whole sequence's inputs in pinned memory (u8 frames as the decoder delivers them)
    g = torch.Generator().manual_seed(7)
    frames_u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8).pin_memory()
    flows_h = syn.smooth_flow(n - 1, h, w, 8.0, seed=1).pin_memory()
    inv_h = syn.inv_depth(n - 1, h, w, seed=2).pin_memory()
    la, lb = (t.to(dev) for t in syn.logits(h, w, seed=3))
    local = [range(c.start - lo, c.stop - lo) for c in mine]
    n_local = sum(len(c) for c in mine)
    out = torch.empty((n_local, s * h, s * w, 3), dtype=torch.uint8, device=dev)

    def one_pass():
        fr = frames_u8[lo:hi].to(dev, non_blocking=True).to(torch.float32)
        fl = flows_h[lo:hi - 1].to(dev, non_blocking=True)
        iv = inv_h[lo:hi - 1].to(dev, non_blocking=True)
        run_sequence(vsr, fr, fl, iv, lambda k: (la, lb), local, out=out)
        return gather_frames_ragged(out)

    # warm-up: three windows (plan build, weight packing, kernel attributes)
    fr = frames_u8[lo:lo + T + 2].to(dev).to(torch.float32)
    run_sequence(vsr, fr, flows_h[lo:lo + T + 1].to(dev), inv_h[lo:lo + T + 1].to(dev), lambda k: (la, lb), [range(0, 3)])
    barrier(world)
    times = []
    allf = None
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(world)
        e0.record()
        allf = one_pass()
        e1.record()
        barrier(world)
        times.append(max_over_ranks(e0.elapsed_time(e1), world, dev))
    ms = sum(times) / len(times)
    n_out = n - T + 1
    assert allf.shape[0] == n_out
    if rank == 0:
        per_rank = [sum(len(c) for c in (shard_chunks if args.exact_chunks else shard_chunks_even)(chunks, world, r))
                    for r in range(world)]
        print(json.dumps({"config": f"C5: {n}-frame 720x360 -> 2880x1440 4x sequence, 3-frame windows (M=8), "
                                    "reference chunking (20 chunks, recurrence reset at chunk starts), sharded; "
                                    "u8 frame all-gather inside the timed region; host (pinned) u8 frames + geometry",
                          "metric": "sr_frames_per_s_4x_sequence", "unit": "frames/s", "n_gpus": world,
                          "value": n_out / (ms / 1e3), "seconds_per_sequence": ms / 1e3, "passes": args.steps,
                          "scaling": "strong", "windows_per_rank": per_rank,
                          "sharding": "whole chunks (identical to the single-GPU result)" if args.exact_chunks
                          else "even by windows, chunks cut at rank boundaries (one extra recurrence reset per cut)",
                          "gathered_bytes": int(allf.numel())}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c3", "c4", "c5", "c2t3"])
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--exact-chunks", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"c3": 20, "c4": 3, "c5": 1, "c2t3": 10}[args.config]
    {"c3": run_c3, "c4": run_c4, "c5": run_c5, "c2t3": run_c2t3}[args.config](args)


if __name__ == "__main__":
    main()
