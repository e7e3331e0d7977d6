"""CPU restatement of the decomposition the x2 strided-conv kernel uses (csrc/srfbn.cu build_downconv2 /
pack_downconv2, csrc/igemm.cuh EPI_DOWN2), checked against torch's Conv2d(32,32,6,2,2) on the same input.

The kernel never forms the 36 taps as shifted input boxes.  With (ky, kx) = (2a + py, 2b + px):
  * K chunk (py, a): the pixel pairs (px, channel) = 64 values of row 2(Y + a - 1) + py of the HR map,
  * N = 96 accumulator columns = column tap b x 32 output channels,
  * P_b[Y, X'] = sum over chunks of pair(Y + a - 1, py, X') . W[b*32 + o, chunk*64 + px*32 + c],
  * out[Y, X] = P_0[Y, X-1] + P_1[Y, X] + P_2[Y, X+1]   (the epilogue's two lane shuffles),
  * 16-column tiles with origin -1 and stride 14 finish their columns 1..14; zero fill outside the image.
This test is the host-side statement of that algebra (SRFBN's k6 s2 p2 geometry; the reference
hard-wires x4, SRProjectionModule.py:101-103), so a change of the packing order or of the tile walk that the GPU
parity tests would catch on the box is also caught here without a GPU."""
import torch
import torch.nn.functional as F


def pack_downconv2(w):
    """(o, c, 6, 6) -> [96 = b*32 + o][384 = (py*3 + a)*64 + px*32 + c], as csrc/srfbn.cu pack_downconv2."""
    out = torch.zeros(96, 384)
    for b in range(3):
        for py in range(2):
            for a in range(3):
                for px in range(2):
                    ky, kx = 2 * a + py, 2 * b + px
                    k0 = (py * 3 + a) * 64 + px * 32
                    out[b * 32:(b + 1) * 32, k0:k0 + 32] = w[:, :, ky, kx]
    return out


def downconv2_by_tiles(x, w, bias):
    """x (H2, W2, 32) with H2 = 2h, W2 = 2w; returns (h, w, 32) the way the kernel's tiles compute it."""
    H2, W2, _ = x.shape
    h, wd = H2 // 2, W2 // 2
    wp = pack_downconv2(w)
    pairs = x.reshape(H2, wd, 64)                       # a row of the parity maps: w pixel pairs of (px, channel)

    def pair_row(r, py, xs):                            # TMA box read: zero fill outside the tensor
        out = torch.zeros(len(xs), 64)
        if 0 <= r < h:
            for i, xx in enumerate(xs):
                if 0 <= xx < wd:
                    out[i] = pairs[2 * r + py, xx]
        return out

    y = torch.full((h, wd, 32), float("nan"))
    for Y in range(h):
        for x0 in range(-1, wd, 14):                    # tile origin -1, stride 14, 16 columns
            xs = list(range(x0, x0 + 16))
            acc = torch.zeros(16, 96)
            for py in range(2):
                for a in range(3):
                    A = pair_row(Y + a - 1, py, xs)     # [16, 64]
                    k0 = (py * 3 + a) * 64
                    acc += A @ wp[:, k0:k0 + 64].t()
            for xi in range(1, 15):                     # columns 1..14 are finished by this tile
                X = x0 + xi
                if X < wd:
                    y[Y, X] = acc[xi - 1, 0:32] + acc[xi, 32:64] + acc[xi + 1, 64:96] + bias
    return y


def test_output_shift_decomposition_equals_conv2d():
    g = torch.Generator().manual_seed(7)
    for h, wd in [(3, 5), (4, 14), (2, 15), (3, 29)]:
        x = torch.randn((2 * h, 2 * wd, 32), generator=g)
        w = torch.randn((32, 32, 6, 6), generator=g) / 30
        b = torch.randn(32, generator=g)
        want = F.conv2d(x.permute(2, 0, 1)[None], w, b, stride=2, padding=2)[0].permute(1, 2, 0)
        got = downconv2_by_tiles(x, w, b)
        assert torch.isfinite(got).all(), "every output pixel is finished by exactly one tile"
        assert (got - want).abs().max().item() < 1e-3
