"""CPU: properties of the C oracle itself (hand-checkable cases)."""
import numpy as np

from oracle import oracle as orc


def test_resample_identity_and_integer_shift():
    rng = np.random.default_rng(0)
    img = rng.random((2, 3, 7, 9)).astype(np.float32)
    zero = np.zeros((2, 2, 7, 9), np.float32)
    assert np.array_equal(orc.resample2d_nchw(img, zero), img)
    flow = zero.copy()
    flow[:, 0] = 2.0   # channel 0 is horizontal (resample2d_kernel.cu:40)
    flow[:, 1] = -1.0
    out = orc.resample2d_nchw(img, flow)
    xs = np.clip(np.arange(9) + 2, 0, 8)
    ys = np.clip(np.arange(7) - 1, 0, 6)
    assert np.array_equal(out, img[:, :, ys][:, :, :, xs])   # border replicate, no zero padding


def test_resample_half_pixel_is_mean_and_nearest_ties_round_up():
    img = np.arange(12, dtype=np.float32).reshape(1, 1, 3, 4)
    flow = np.zeros((1, 2, 3, 4), np.float32)
    flow[:, 0] = 0.5
    out = orc.resample2d_nchw(img, flow)
    assert np.allclose(out[0, 0, :, :3], img[0, 0, :, :3] + 0.5)
    assert np.array_equal(out[0, 0, :, 3], img[0, 0, :, 3])  # xL = xR = W-1 at the border
    near = orc.resample2d_nchw(img, flow, bilinear=False)
    assert np.array_equal(near[0, 0, :, :3], img[0, 0, :, 1:])  # floor(x + 0.5 + 0.5) = x + 1


def test_nhwc_and_nchw_agree_bitwise():
    rng = np.random.default_rng(1)
    img = (rng.random((2, 11, 13, 3)) * 255).astype(np.float32)
    flow = ((rng.random((2, 11, 13, 2)) - 0.5) * 9).astype(np.float32)
    a = orc.warp_nhwc(img, flow)
    b = orc.resample2d_nchw(img.transpose(0, 3, 1, 2), flow.transpose(0, 3, 1, 2))
    assert np.array_equal(a, b.transpose(0, 2, 3, 1))
    assert np.array_equal(orc.warp_nhwc(img, flow, threads=3), a)


def test_label_warp_matches_nearest_resample():
    rng = np.random.default_rng(2)
    lab = (rng.random((2, 9, 10)) > 0.6).astype(np.uint8)
    flow = ((rng.random((2, 9, 10, 2)) - 0.5) * 7).astype(np.float32)
    a = orc.warp_labels(lab, flow)
    b = orc.warp_nhwc(lab[..., None].astype(np.float32), flow, bilinear=False)[..., 0]
    assert np.array_equal(a, b.astype(np.uint8))


def test_channelnorm():
    x = np.array([3.0, 4.0], np.float32).reshape(1, 2, 1, 1)
    assert orc.channelnorm_nchw(x)[0, 0, 0, 0] == 5.0
    rng = np.random.default_rng(3)
    y = rng.random((2, 5, 4, 6)).astype(np.float32)
    assert np.allclose(orc.channelnorm_nchw(y)[:, 0], np.sqrt((y.astype(np.float64) ** 2).sum(1)), rtol=1e-6)
    assert np.array_equal(orc.channelnorm_nhwc(y.transpose(0, 2, 3, 1)), orc.channelnorm_nchw(y)[:, 0])


def test_flow_projection_zero_flow():
    h, w = 5, 6
    flow = np.zeros((1, h, w, 2), np.float32)
    proj, wsum, count, hole = orc.flow_projection(flow)
    # each pixel splats to itself, right, below, below-right (clamped duplicates at the borders)
    expect = np.zeros((h, w), np.int32)
    for y in range(h):
        for x in range(w):
            for yy in (y, min(y + 1, h - 1)):
                for xx in (x, min(x + 1, w - 1)):
                    expect[yy, xx] += 1
    assert np.array_equal(count[0], expect)
    assert count.sum() == 4 * h * w
    assert not hole.any() and not proj.any()
    assert np.array_equal(wsum[0], expect.astype(np.float32))


def test_flow_projection_constant_shift_holes_and_fill():
    h, w = 6, 8
    flow = np.zeros((1, h, w, 2), np.float32)
    flow[..., 0] = 3.0
    proj, wsum, count, hole = orc.flow_projection(flow)
    assert hole[0, :, :3].all() and not hole[0, :, 3:].any()      # trailing band is all holes
    assert np.allclose(proj[0, :, 3:, 0], -3.0) and np.allclose(proj[0, :, 3:, 1], 0.0)
    assert np.allclose(proj[0, :, :3, 0], -3.0)                    # filled from the right neighbour
    assert (count[0][hole[0] == 1] == 0).all()


def test_flow_projection_all_out_of_range_stays_zero():
    flow = np.full((1, 4, 4, 2), 100.0, np.float32)
    proj, wsum, count, hole = orc.flow_projection(flow)
    assert hole.all() and not count.any() and not proj.any() and not wsum.any()


def test_depth_projection_nearer_surface_dominates():
    h, w = 1, 6
    flow = np.zeros((1, h, w, 2), np.float32)
    inv = np.full((1, h, w), 0.1, np.float32)
    flow[0, 0, 0, 0] = 3.0     # pixel 0 (near, inv depth 1) lands on pixel 3 (far, static)
    inv[0, 0, 0] = 1.0
    proj, wsum, count, hole = orc.flow_projection(flow, inv)
    plain, _, _, _ = orc.flow_projection(flow)
    assert proj[0, 0, 3, 0] < plain[0, 0, 3, 0] < 0    # the weighted estimate sits nearer to -3
    # h == 1: yB == yT, so every contribution lands twice (clamped duplicate target, App. B step 2)
    assert np.isclose(wsum[0, 0, 3], 2 * (1.0 + 0.1 + 0.1))
    assert np.array_equal(count, orc.flow_projection(flow)[2])   # counts do not depend on depth


def test_flow_projection_nan_flow_is_skipped():
    flow = np.zeros((1, 3, 3, 2), np.float32)
    flow[0, 1, 1, 0] = np.nan
    _, _, count, _ = orc.flow_projection(flow)
    assert count.sum() == 4 * 8


def test_backward_oracles_against_torch_autograd():
    """CPU: the backward restatements vs torch autograd of an equivalent differentiable formulation
    (positive coordinates only, where the reference's int() truncation equals floor)."""
    import torch
    from oracle import oracle as orc
    g = torch.Generator().manual_seed(0)
    B, C, H, W = 1, 3, 12, 16
    img = torch.rand((B, C, H, W), generator=g, dtype=torch.float64)
    flow = torch.rand((B, 2, H, W), generator=g, dtype=torch.float64) * 1.5 + 0.1     # keeps x+dx > 0
    flow[:, 0, :, -3:] = 0.25
    flow[:, 1, -3:, :] = 0.25                                                          # stay inside: no clamping
    gout = torch.randn((B, C, H, W), generator=g, dtype=torch.float64)
    imgt = img.clone().requires_grad_(True)
    ft = flow.clone().requires_grad_(True)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    xf, yf = xs + ft[:, 0], ys + ft[:, 1]
    x0, y0 = xf.detach().floor(), yf.detach().floor()
    a, b = xf - x0, yf - y0
    x0i, y0i = x0.long().clamp(0, W - 1), y0.long().clamp(0, H - 1)
    x1i, y1i = (x0i + 1).clamp(0, W - 1), (y0i + 1).clamp(0, H - 1)

    def tap(yy, xx):
        return imgt[0][:, yy[0], xx[0]]

    out = (1 - a) * (1 - b) * tap(y0i, x0i) + a * (1 - b) * tap(y0i, x1i) + (1 - a) * b * tap(y1i, x0i) + a * b * tap(y1i, x1i)
    (out * gout[0]).sum().backward()
    g1, g2 = orc.resample2d_backward(img.float().numpy(), flow.float().numpy(), gout.float().numpy())
    assert np.abs(g1 - imgt.grad.numpy()).max() < 1e-4
    assert np.abs(g2 - ft.grad.numpy()).max() < 1e-4
    x = torch.randn((2, 3, 5, 7), generator=g, dtype=torch.float64).requires_grad_(True)
    n = (x * x).sum(1, keepdim=True).sqrt()
    go = torch.randn(n.shape, generator=g, dtype=torch.float64)
    (n * go).sum().backward()
    got = orc.channelnorm_backward(x.detach().float().numpy(), n.detach().float().numpy(), go.float().numpy())
    assert np.abs(got - x.grad.numpy()).max() < 1e-5


def test_correlation_oracle_against_direct_numpy():
    """CPU: or_correlation vs a direct numpy evaluation of the definition (correlation_cuda_kernel.cu:73-147)."""
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    B, C, H, W, pad, k, md, s1, s2 = 1, 5, 9, 11, 4, 1, 4, 1, 2
    a = rng.standard_normal((B, C, H, W)).astype(np.float32)
    b = rng.standard_normal((B, C, H, W)).astype(np.float32)
    got = orc.correlation(a, b, pad, k, md, s1, s2)
    R = md // s2
    D = 2 * R + 1
    assert got.shape == (B, D * D, H, W)
    ap = np.pad(a, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    bp = np.pad(b, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    for tj in range(-R, R + 1):
        for ti in range(-R, R + 1):
            ref = np.zeros((H, W))
            for y in range(H):
                for x in range(W):
                    y1, x1 = y + md, x + md
                    ref[y, x] = (ap[0, :, y1, x1].astype(np.float64) * bp[0, :, y1 + tj * s2, x1 + ti * s2]).sum() / C
            assert np.abs(got[0, (tj + R) * D + ti + R] - ref).max() < 1e-5
