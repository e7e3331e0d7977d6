"""A SECOND, independent restatement of the splat contract, written from the text of the
appendix -- four targets per source, `index_put_(accumulate=True)`, fill by explicit search loops -- and checked against
the C oracle (oracle/oracle.c::or_flow_projection, which factors nothing either but comes from another hand-pass over
the same text) and, on the GPU, against the product kernels (which DO factor the splat into cells + a 2x2 box sum).

The reference has no forward splat at all (SURVEY.md banner), so a1/a2 stay "parity unpinned": two restatements and one
kernel agreeing is the most this row can get, and it is what would catch a transcription error such as int() vs floor
or the multiplicity of the clamped duplicate targets."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from video_super_resolution_b200 import synthetic as syn


def appendix_b(flow: torch.Tensor, inv_depth: torch.Tensor | None):
    """flow (h,w,2), inv_depth (h,w)|None -> proj (h,w,2) f32, wsum (h,w) f32, count (h,w) i32, hole (h,w) u8."""
    h, w, _ = flow.shape
    fx, fy = flow[..., 0], flow[..., 1]
    D = torch.ones((h, w)) if inv_depth is None else inv_depth
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    x2, y2 = xs + fx, ys + fy                                              # step 2, fp32
    ok = (x2 >= 0) & (x2 <= w - 1) & (y2 >= 0) & (y2 <= h - 1)             # NaN fails every comparison
    x2, y2, fxs, fys, Ds = x2[ok], y2[ok], fx[ok], fy[ok], D[ok]
    xL, yT = x2.to(torch.int64), y2.to(torch.int64)                        # int(): truncation (x2, y2 >= 0 here)
    xR, yB = torch.clamp(xL + 1, max=w - 1), torch.clamp(yT + 1, max=h - 1)
    sx = torch.zeros((h, w), dtype=torch.float64)
    sy = torch.zeros((h, w), dtype=torch.float64)
    ws = torch.zeros((h, w), dtype=torch.float64)
    cnt = torch.zeros((h, w), dtype=torch.int64)
    cx = (-fxs * Ds).double()                                              # the contribution is an fp32 product
    cy = (-fys * Ds).double()
    one = torch.ones_like(xL)
    for ty, tx in ((yT, xL), (yT, xR), (yB, xL), (yB, xR)):                # a clamped duplicate target is hit twice
        sx.index_put_((ty, tx), cx, accumulate=True)
        sy.index_put_((ty, tx), cy, accumulate=True)
        ws.index_put_((ty, tx), Ds.double(), accumulate=True)
        cnt.index_put_((ty, tx), one, accumulate=True)
    hit = cnt > 0                                                          # step 3 / 4 (inverse depths > 0 here)
    norm = torch.zeros((h, w, 2), dtype=torch.float32)
    wsf = ws.float()
    norm[..., 0][hit] = sx.float()[hit] / wsf[hit]
    norm[..., 1][hit] = sy.float()[hit] / wsf[hit]
    out = norm.clone()
    hit_np = hit.numpy()
    for y in range(h):                                                     # step 4: explicit searches, hole pixels only
        for x in np.nonzero(~hit_np[y])[0]:
            found = []
            for xx in range(x - 1, -1, -1):
                if hit_np[y, xx]:
                    found.append(norm[y, xx]); break
            for xx in range(x + 1, w):
                if hit_np[y, xx]:
                    found.append(norm[y, xx]); break
            for yy in range(y - 1, -1, -1):
                if hit_np[yy, x]:
                    found.append(norm[yy, x]); break
            for yy in range(y + 1, h):
                if hit_np[yy, x]:
                    found.append(norm[yy, x]); break
            if found:
                acc = torch.zeros(2)
                for v in found:                                            # left, right, up, down
                    acc = acc + v
                out[y, x] = acc / float(len(found))
    wsum = torch.where(hit, wsf, torch.zeros_like(wsf))
    return out.numpy(), wsum.numpy(), cnt.to(torch.int32).numpy(), (~hit).to(torch.uint8).numpy()


def _cases():
    yield "smooth8", syn.smooth_flow(1, 40, 56, 8.0, seed=1)[0], syn.inv_depth(1, 40, 56, seed=2)[0]
    yield "iid64", syn.random_flow(1, 48, 80, 64.0, seed=3)[0], syn.inv_depth(1, 48, 80, seed=4)[0]
    f, d = syn.occlusion_scene(1, 40, 120, shift=30.0, seed=5)
    yield "occlusion", f[0] + syn.random_flow(1, 40, 120, 0.75, seed=6)[0], d[0]
    yield "unweighted", syn.random_flow(1, 33, 47, 5.0, seed=7)[0], None
    g = torch.Generator().manual_seed(8)
    f = torch.randint(-3, 4, (17, 33, 2), generator=g).float() + torch.randint(0, 2, (17, 33, 2), generator=g).float() * 0.5
    f[:, -1, 0] = 0.0                 # last column / last row stay on themselves: the clamped duplicate targets
    f[-1, :, 1] = 0.0
    f[0, 0] = torch.tensor([32.0, 16.0])          # corner source -> opposite corner exactly
    f[3, 4, 0] = float("nan")
    yield "border_hits_nan", f, syn.inv_depth(1, 17, 33, seed=9)[0]
    yield "one_row", syn.random_flow(1, 1, 40, 6.0, seed=10)[0], None
    yield "one_col", syn.random_flow(1, 40, 1, 6.0, seed=11)[0], syn.inv_depth(1, 40, 1, seed=12)[0]


@pytest.mark.parametrize("name,flow,inv", list(_cases()), ids=[c[0] for c in _cases()])
def test_c_oracle_agrees_with_the_independent_restatement(name, flow, inv):
    proj, wsum, count, hole = appendix_b(flow, inv)
    o_proj, o_wsum, o_count, o_hole = orc.flow_projection(flow.numpy()[None], None if inv is None else inv.numpy()[None])
    assert np.array_equal(o_count[0], count)
    assert np.array_equal(o_hole[0], hole)
    assert np.abs(o_proj[0] - proj).max() <= 1e-4
    assert np.abs(o_wsum[0] - wsum).max() <= 1e-4 * max(1.0, float(wsum.max()))
    if name in ("iid64", "occlusion"):
        assert hole.any() and count.max() >= 4            # the case exercises collisions and the fill


@pytest.mark.gpu
@pytest.mark.parametrize("name,flow,inv", list(_cases()), ids=[c[0] for c in _cases()])
def test_kernels_agree_with_the_independent_restatement(name, flow, inv):
    from video_super_resolution_b200 import ops
    proj, wsum, count, hole = appendix_b(flow, inv)
    for md in (None, 16.0):          # general path; tile path (or its device-side fallback when |flow| > 16)
        g = ops.project_flow(flow[None].cuda(), None if inv is None else inv[None].cuda(), md)
        assert np.array_equal(g[2][0].cpu().numpy(), count), md
        assert np.array_equal(g[3][0].cpu().numpy(), hole), md
        assert np.abs(g[0][0].cpu().numpy() - proj).max() <= 1e-3, md
        assert np.abs(g[1][0].cpu().numpy() - wsum).max() <= 1e-3 * max(1.0, float(wsum.max())), md
