"""GPU parity of the whole per-window hot path (WarpFusePipeline / VSR.forward_geometry) against the
CPU oracle: integer outputs bit-exact, fp32 front within 1e-3, SR frame within the BF16 bound."""
import math

import numpy as np
import pytest
import torch

from oracle import oracle as orc
from oracle import srfbn_oracle as so
from video_super_resolution_b200 import ops, synthetic as syn
from video_super_resolution_b200.my_packages.SRProjection.SRProjectionModule import SRProjectionModule
from video_super_resolution_b200.network.video_super_resolution import VSR
from video_super_resolution_b200.pipeline import WarpFusePipeline

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(T, h, w, seed):
    la, lb = syn.logits(h, w, seed=seed + 3)
    return (syn.frames(T, h, w, seed=seed), syn.smooth_flow(T - 1, h, w, 5.0, seed=seed + 1),
            syn.inv_depth(T - 1, h, w, seed=seed + 2), la, lb)


@pytest.mark.parametrize("T,h,w,max_disp", [(3, 32, 48, None), (5, 21, 37, 5.0), (7, 16, 24, None), (7, 70, 200, 5.0),
                                             (2, 20, 30, 5.0)])
def test_front_matches_oracle(T, h, w, max_disp):
    fr, fl, d, la, lb = _inputs(T, h, w, seed=T)
    pipe = WarpFusePipeline(T, h, w, SRProjectionModule(num_maps=3 * T - 1), 4, device=DEV, run_fusion=False,
                            max_disp=max_disp)
    est = torch.rand((3, h, w), generator=torch.Generator().manual_seed(9)) * 255
    r = pipe.project_and_warp(fr.to(DEV), fl.to(DEV), d.to(DEV), la.to(DEV), lb.to(DEV), est.to(DEV))
    stack, want = orc.warp_fuse_front(fr.numpy(), fl.numpy(), d.numpy(), la.numpy(), lb.numpy(), est.numpy())
    for k in ("count_flow", "hole_flow", "count_depth", "hole_depth", "mask"):
        assert np.array_equal(r[k].cpu().numpy(), want[k]), k                  # bit-exact integer outputs
    # the label warp is bit-exact GIVEN its flow (checked with the oracle's flow at the end); with the pipeline's own
    # projected flow (fp32 sums in another order, <= 1e-3 px) a label may differ only where the sampling position is
    # that close to a rounding tie of floor(v + 0.5)
    diff = np.argwhere(r["mask_warped"].cpu().numpy() != want["mask_warped"])
    if T > 1 and len(diff):
        cf = want["centre_flows"][T // 2 - 1]
        for y, x in diff:
            fx, fy = x + cf[y, x, 0], y + cf[y, x, 1]
            assert min(abs(fx + 0.5 - round(fx + 0.5)), abs(fy + 0.5 - round(fy + 0.5))) <= 1e-3 * T, (y, x, fx, fy)
    assert len(diff) <= 2
    for k in ("proj_flow", "proj_depth", "wsum"):
        assert np.abs(r[k].cpu().numpy() - want[k]).max() <= 1e-3, k           # fp32 sums in atomic order
    # the chained centre -> neighbour flows: sums of (T-1)/2 sampled projected flows, each within 1e-3
    assert np.abs(r["centre_flows"].cpu().numpy() - want["centre_flows"]).max() <= 1e-3 * T
    # the warp is bit-exact GIVEN its flow; here the flow itself differs by atomic ordering (<=1e-3 px),
    # which moves a 0..255 image by at most |gradient| * 1e-3
    assert np.abs(r["warped"].cpu().numpy() - want["warped"]).max() <= 0.5
    got = pipe.stack.cpu().numpy()
    assert got.shape == stack.shape
    assert np.abs(got - stack).max() <= 0.5
    assert np.array_equal(got[3 * T - 2], est.numpy())                         # estimate slot copied verbatim
    # ... and held to the north star's 1e-3 with the flow taken out of the comparison: the pipeline's own warp
    # kernel driven by the ORACLE's centre flows against the oracle's warp (reference arithmetic, Appendix A)
    c = T // 2
    warped2, resid2 = ops.warp_window(fr.to(DEV), torch.from_numpy(want["centre_flows"]).to(DEV), c, 2)
    assert np.abs(warped2.cpu().numpy() - want["warped"]).max() <= 1e-3
    assert np.abs(resid2.cpu().numpy() - want["resid"]).max() <= 2e-3
    exact, _ = ops.warp_window(fr.to(DEV), torch.from_numpy(want["centre_flows"]).to(DEV), c, 1)
    assert np.array_equal(exact.cpu().numpy(), want["warped"])                 # exact mode: bit for bit
    lab = ops.warp_labels(r["mask"].unsqueeze(0), torch.from_numpy(want["centre_flows"][c - 1:c]).to(DEV))[0]
    assert np.array_equal(lab.cpu().numpy(), want["mask_warped"])


def test_estimate_fallback_is_lr_frame_0():
    """No estimate (first window of a chunk / shard): the slot holds LR frame 0 of the window, as the reference's
    `else data_clone[0:1]` (network/video_super_resolution.py:37-38)."""
    T, h, w = 5, 18, 26
    fr, fl, d, la, lb = _inputs(T, h, w, seed=3)
    pipe = WarpFusePipeline(T, h, w, SRProjectionModule(num_maps=3 * T - 1), 4, device=DEV, run_fusion=False)
    pipe.project_and_warp(fr.to(DEV), fl.to(DEV), d.to(DEV), la.to(DEV), lb.to(DEV), None)
    assert np.array_equal(pipe.stack[3 * T - 2].cpu().numpy(), fr[0].permute(2, 0, 1).numpy())
    hr = torch.rand((3, 4 * h, 4 * w), generator=torch.Generator().manual_seed(2)) * 255
    pipe.project_and_warp(fr.to(DEV), fl.to(DEV), d.to(DEV), la.to(DEV), lb.to(DEV), None, estimate_hr=hr.to(DEV))
    assert np.array_equal(pipe.stack[3 * T - 2].cpu().numpy(), hr[:, ::4, ::4].numpy())     # :35-37 nearest downsize


@pytest.mark.parametrize("T", [2, 3, 5, 7])
def test_translating_scene_is_aligned_to_the_centre_frame(T):
    """A scene that translates by a constant integer velocity: frame t = base shifted by t*v, every flow n -> n+1
    is the constant v.  Every neighbour -- before AND after the centre -- must come out equal to the centre frame
    (away from the borders), the residual maps must vanish there, and a mask defined in frame c-1 must land on the
    centre frame's pixels."""
    h, w, vx, vy = 48, 64, 3, -2
    g = torch.Generator().manual_seed(T)
    big = torch.rand((h + 2 * 8 * T, w + 2 * 8 * T, 3), generator=g) * 255
    o = 8 * T
    # content at p in frame t sits at p + v in frame t+1  <=>  frame_t(p) = base(p - t*v)
    frames = torch.stack([big[o - t * vy:o - t * vy + h, o - t * vx:o - t * vx + w] for t in range(T)]).contiguous()
    flows = torch.zeros((T - 1, h, w, 2))
    flows[..., 0], flows[..., 1] = float(vx), float(vy)
    inv = syn.inv_depth(T - 1, h, w, seed=1)
    c = T // 2
    la = torch.full((h, w), -20.0)
    la[10:20, 12:30] = 20.0                                   # mask (frame c-1 coordinates): a rectangle
    for md in (None, 3.0):
        pipe = WarpFusePipeline(T, h, w, SRProjectionModule(num_maps=3 * T - 1), 4, device=DEV, run_fusion=False,
                                max_disp=md)
        r = pipe.project_and_warp(frames.to(DEV), flows.to(DEV), inv.to(DEV), la.to(DEV), la.to(DEV))
        m = 3 * T + 2                                         # border margin: chain length x |v| + taps
        centre = frames[c, m:h - m, m:w - m]
        for n, t in enumerate([t for t in range(T) if t != c]):
            got = r["warped"][n].cpu()[m:h - m, m:w - m]
            assert (got - centre).abs().max().item() <= 1e-3, (T, t)
            assert r["resid"][n].cpu()[m:h - m, m:w - m].abs().max().item() <= 2e-3, (T, t)
        want_mask = torch.zeros((h, w), dtype=torch.uint8)
        want_mask[10 + vy:20 + vy, 12 + vx:30 + vx] = 1       # the rectangle moved by one frame of motion
        assert torch.equal(r["mask_warped"].cpu()[m:h - m, m:w - m], want_mask[m:h - m, m:w - m])


def test_estimate_slot_bit_exact():
    h, w = 19, 23
    g = torch.Generator().manual_seed(1)
    hr = torch.rand((3, 4 * h, 4 * w), generator=g) * 255
    mask = (torch.rand((h, w), generator=g) > 0.6).to(torch.uint8)
    slot = torch.empty((3, h, w), device=DEV)
    ops.estimate_slot(hr.to(DEV), mask.to(DEV), slot, 4)
    assert np.array_equal(slot.cpu().numpy(), orc.estimate_slot(hr.numpy(), mask.numpy(), 4))
    ops.estimate_slot(hr.to(DEV), None, slot, 4)
    assert np.array_equal(slot.cpu().numpy(), orc.estimate_slot(hr.numpy(), None, 4))


def test_vsr_forward_geometry_against_oracle():
    T, h, w = 3, 24, 32
    M = 3 * T - 1
    fr, fl, d, la, lb = _inputs(T, h, w, seed=11)
    sd = so.init_state_dict(num_maps=M, seed=2, gain=2.3)
    vsr = VSR(window=T)
    vsr.model.load_state_dict(sd)
    out = vsr.forward_geometry(fr.to(DEV), fl.to(DEV), d.to(DEV), la.to(DEV), lb.to(DEV), None)
    assert tuple(out.shape) == (1, 4 * h, 4 * w, 3)
    stack, want = orc.warp_fuse_front(fr.numpy(), fl.numpy(), d.numpy(), la.numpy(), lb.numpy(), None)
    x = torch.from_numpy(stack)
    out1 = so.forward(x, sd)
    x[M - 1] = torch.from_numpy(orc.estimate_slot(out1[0].numpy(), want["mask_warped"], 4))
    ref = so.forward(x, sd)[0].permute(1, 2, 0)
    err = (out[0].cpu() - ref).abs()
    scale = ref.abs().mean().item() + 1.0
    assert torch.isfinite(out).all()
    assert err.max().item() <= 0.05 * scale + 0.5, (err.max().item(), scale)
    mse = (err ** 2).mean().item()
    assert 10 * math.log10(max(ref.abs().max().item(), 1.0) ** 2 / max(mse, 1e-20)) > 40.0


def test_step_u8_frame_from_the_fuse_kernel():
    """The u8 frame the fc-fuse kernel writes equals clamp(0,255) -> round half to even -> u8 of the fp32 frame
    (pipeline.quantise_u8, the loader's pixel format), with and without the fp32 output."""
    from video_super_resolution_b200.pipeline import quantise_u8
    T, h, w = 3, 20, 28
    M = 3 * T - 1
    fr, fl, d, la, lb = _inputs(T, h, w, seed=5)
    sr = SRProjectionModule(num_maps=M)
    sr.load_state_dict(so.init_state_dict(num_maps=M, seed=1, gain=2.3))
    pipe = WarpFusePipeline(T, h, w, sr, 4, device=DEV)
    args = [t.to(DEV) for t in (fr, fl, d, la, lb)]
    u8 = torch.zeros((4 * h, 4 * w, 3), dtype=torch.uint8, device=DEV)
    y = pipe.step(*args, out_u8=u8)
    assert torch.equal(u8, quantise_u8(y)) and int(u8.max()) > 0
    u8b = torch.zeros_like(u8)
    r = pipe.step(*args, out_u8=u8b, want_f32=False)
    assert r is u8b and torch.equal(u8b, u8)


@pytest.mark.parametrize("T", [3, 5])
def test_step_with_reuse_of_the_unchanged_maps_is_bit_identical(T):
    """reuse='frames' / 'unchanged': the fuse pass convolves only the maps that changed since pass 1 and takes the
    per-map images of the others from pass 1 -- same frame, bit for bit, eagerly and from a captured graph."""
    from video_super_resolution_b200.pipeline import GraphedStep
    h, w = 22, 36
    M = 3 * T - 1
    sr = SRProjectionModule(num_maps=M)
    sr.load_state_dict(so.init_state_dict(num_maps=M, seed=2, gain=2.3))
    base = WarpFusePipeline(T, h, w, sr, 4, device=DEV, max_disp=5.0)
    for mode in ("frames", "unchanged"):
        sr2 = SRProjectionModule(num_maps=M)
        sr2.load_state_dict(so.init_state_dict(num_maps=M, seed=2, gain=2.3))
        pipe = WarpFusePipeline(T, h, w, sr2, 4, device=DEV, max_disp=5.0, reuse=mode)
        g = None
        for seed in (7, 8, 9):
            args = [t.to(DEV) for t in _inputs(T, h, w, seed=seed)]
            u8a = torch.zeros((4 * h, 4 * w, 3), dtype=torch.uint8, device=DEV)
            u8b = torch.zeros_like(u8a)
            want = base.step(*args, out_u8=u8a)
            got = pipe.step(*args, out_u8=u8b)
            assert torch.equal(got, want) and torch.equal(u8a, u8b), (mode, seed)
            if g is None:
                g = GraphedStep(pipe, want_f32=True)
            g.load(*args)
            assert torch.equal(g.replay(), want), (mode, seed)
    with pytest.raises(ValueError):
        WarpFusePipeline(T, h, w, sr, 4, device=DEV, reuse="all")


def test_graphed_step_replays_the_same_frame():
    """The whole per-frame sequence captured into one CUDA graph: replays reproduce the eager frames bit for bit, for
    changing inputs, on both projection paths (the gated cooperative fallback launch is part of the graph)."""
    from video_super_resolution_b200.pipeline import GraphedStep
    T, h, w = 5, 24, 40
    M = 3 * T - 1
    sr = SRProjectionModule(num_maps=M)
    sr.load_state_dict(so.init_state_dict(num_maps=M, seed=1, gain=2.3))
    for md in (None, 5.0):
        pipe = WarpFusePipeline(T, h, w, sr, 4, device=DEV, max_disp=md)
        g = GraphedStep(pipe, want_f32=True)
        for seed in (3, 4, 5):
            args = [t.to(DEV) for t in _inputs(T, h, w, seed=seed)]
            u8 = torch.zeros((4 * h, 4 * w, 3), dtype=torch.uint8, device=DEV)
            want = pipe.step(*args, out_u8=u8).clone()
            g.load(*args)
            got = g.replay()
            assert torch.equal(got, want) and torch.equal(g.frame_u8, u8), (md, seed)
    from video_super_resolution_b200 import _lib
    _lib.launch_count_reset()
    g.replay()
    torch.cuda.synchronize()
    assert _lib.launch_count() == 0          # a replay issues no launch calls of its own


def test_module_surfaces_refuse_missing_estimators():
    from video_super_resolution_b200.my_packages.FlowProjection.FlowProjectionModule import FlowProjectionModule
    m = FlowProjectionModule()
    with pytest.raises(RuntimeError, match="estimator"):
        m(torch.zeros(64, 64, 3, device=DEV), torch.zeros(64, 64, 3, device=DEV))
    flow = syn.smooth_flow(1, 64, 64, 3.0)[0].to(DEV)
    m2 = FlowProjectionModule(estimator=lambda a, b: flow)
    out = m2(torch.zeros(64, 64, 3, device=DEV), torch.zeros(64, 64, 3, device=DEV))
    assert tuple(out.shape) == (64, 64, 3)                       # (h',w',3) like the reference (:33)
    want = orc.flow_projection(flow.cpu().numpy()[None])
    assert np.array_equal(out[..., 2].cpu().numpy().astype(np.uint8), want[3][0])


def test_run_sequence_chunked_and_sharded():
    """The video loop (main.py:190-203) over the reference's chunks: sharding by whole chunks (shard_chunks)
    reproduces the single-run frames bit for bit; the even split differs only inside the chunk it cuts."""
    from video_super_resolution_b200.network.video_super_resolution import VSR
    from video_super_resolution_b200.pipeline import run_sequence, shard_chunks, shard_chunks_even
    from video_super_resolution_b200.utils.video_utils import chunk_windows
    T, h, w, n = 3, 16, 24, 23
    frames = syn.frames(n, h, w, seed=1).to(DEV)
    flows = syn.smooth_flow(n - 1, h, w, 2.0, seed=2).to(DEV)
    inv = syn.inv_depth(n - 1, h, w, seed=3).to(DEV)
    la, lb = (t.to(DEV) for t in syn.logits(h, w, seed=4))
    torch.manual_seed(0)
    vsr = VSR(window=T)
    chunks = chunk_windows(n, T, splitvideonum=4)               # 5-window chunks: [5,5,5,5,1]
    assert [len(c) for c in chunks] == [5, 5, 5, 5, 1]
    whole, idx = run_sequence(vsr, frames, flows, inv, lambda k: (la, lb), chunks)
    assert idx == list(range(n - T + 1)) and whole.shape == (n - T + 1, 4 * h, 4 * w, 3) and whole.dtype == torch.uint8
    parts = [run_sequence(vsr, frames, flows, inv, lambda k: (la, lb), shard_chunks(chunks, 2, r)) for r in range(2)]
    assert parts[0][1] + parts[1][1] == idx
    assert torch.equal(torch.cat([parts[0][0], parts[1][0]]), whole)
    even = [run_sequence(vsr, frames, flows, inv, lambda k: (la, lb), shard_chunks_even(chunks, 2, r)) for r in range(2)]
    cat = torch.cat([even[0][0], even[1][0]])
    cut = even[1][1][0]                                         # first window of rank 1: an extra reset
    assert cut % 5 != 0                                         # ... inside a chunk
    same = [bool(torch.equal(cat[k], whole[k])) for k in idx]
    chunk_end = (cut // 5 + 1) * 5
    assert all(same[:cut]) and all(same[chunk_end:])
    assert not same[cut]


def test_pinned_frame_ring_overlapped_h2d():
    from video_super_resolution_b200.utils.video_utils import PinnedFrameRing
    ring = PinnedFrameRing((3, 8, 12, 3), depth=2, device=DEV)
    outs = []
    for k in range(5):
        src = torch.full((3, 8, 12, 3), float(k))
        dev, ready, slot = ring.put(src)
        torch.cuda.current_stream().wait_event(ready)
        outs.append(dev.sum().clone())
        ring.release(slot)
    torch.cuda.synchronize()
    assert [float(o) for o in outs] == [k * 3 * 8 * 12 * 3 for k in range(5)]
