// correlation.cu -- FlowNetC cost volume (not on the warp-and-fuse hot path).
//   ref: correlation_package/correlation_cuda.cc:10-86 (output geometry, zero padding),
//        correlation_cuda_kernel.cu:46-147 (channels_first repack + correlation_forward),
//        correlation.py:7-30; FlowNetC use: pad 20, kernel 1, max_displacement 20, strides 1 / 2
//        (FlowNetC.py:22) -> 441 output channels.
//
//   out[n, (tj+R)*D + (ti+R), y, x] = 1/(k*k*C) * sum_{j,i in kernel} sum_c
//         pad(in1)[n, c, y1+j, x1+i] * pad(in2)[n, c, y1+j + tj*s2, x1+i + ti*s2]
//   y1 = y*s1 + max_displacement (padded coordinates), R = max_displacement / s2, D = 2R+1.
//
// The reference repacks both inputs into zero-padded NHWC copies and then spends one 32-thread block
// per output PIXEL, re-reading the C channels of both images for each of the D*D displacements
// through a warp-shuffle reduction.  Here a warp owns 32 consecutive output x of one (n, y, tj):
// lane = x keeps the D accumulators of its pixel in registers, in1 is read once per channel and the D
// shifted in2 values are coalesced, L1-resident loads; no repack, no padding copy, no reduction.
// fp32 FMA accumulation in channel order (the reference sums 32 strided partials and tree-reduces,
// so results agree to fp32 rounding, not bit for bit).
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxD = 21;   // displacements per axis kept in registers (FlowNetC: 21)

__global__ void __launch_bounds__(kThreads)
correlation_forward_kernel(const float* __restrict__ in1, const float* __restrict__ in2, float* __restrict__ out,
                           int B, int C, int H, int W, int outH, int outW, int pad, int ksize, int maxdisp, int s1, int s2,
                           int R, int D, int segs) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t n_tasks = (int64_t)B * outH * D * segs;
  const int64_t n_warps = ((int64_t)gridDim.x * kThreads) >> 5;
  const int krad = (ksize - 1) / 2;
  const int64_t HW = (int64_t)H * W;
  for (int64_t task = warp_global; task < n_tasks; task += n_warps) {
    const int seg = (int)(task % segs);
    int64_t r = task / segs;
    const int tjw = (int)(r % D);      // tj + R
    r /= D;
    const int y = (int)(r % outH);
    const int n = (int)(r / outH);
    const int x = seg * 32 + lane;
    // unpadded coordinates of the two window centres
    const int ya = y * s1 + maxdisp - pad, xa = x * s1 + maxdisp - pad;
    const int yb = ya + (tjw - R) * s2;
    float acc[kMaxD];
#pragma unroll
    for (int t = 0; t < kMaxD; ++t) acc[t] = 0.0f;
    const float* p1 = in1 + (int64_t)n * C * HW;
    const float* p2 = in2 + (int64_t)n * C * HW;
    for (int j = -krad; j <= krad; ++j) {
      const int y1 = ya + j, y2 = yb + j;
      const bool row1 = (y1 >= 0) && (y1 < H), row2 = (y2 >= 0) && (y2 < H);
      if (!(row1 && row2)) continue;                 // a zero-padded row on either side contributes nothing
      for (int i = -krad; i <= krad; ++i) {
        const int x1 = xa + i;
        const bool in_a = (x < outW) && (x1 >= 0) && (x1 < W);
        const float* a_ptr = p1 + (int64_t)y1 * W + (in_a ? x1 : 0);
        const float* b_row = p2 + (int64_t)y2 * W;
        for (int c = 0; c < C; ++c) {
          const float a = in_a ? __ldg(a_ptr + (int64_t)c * HW) : 0.0f;
          const float* b_c = b_row + (int64_t)c * HW;
#pragma unroll
          for (int t = 0; t < kMaxD; ++t) {
            if (t < D) {
              const int x2 = x1 + (t - R) * s2;
              const float b = (in_a && x2 >= 0 && x2 < W) ? __ldg(b_c + x2) : 0.0f;
              acc[t] = fmaf(a, b, acc[t]);
            }
          }
        }
      }
    }
    if (x < outW) {
      const float nelems = (float)(ksize * ksize * C);
      float* o = out + (((int64_t)n * D * D + (int64_t)tjw * D) * outH + y) * outW + x;
#pragma unroll
      for (int t = 0; t < kMaxD; ++t)
        if (t < D) o[(int64_t)t * outH * outW] = __fdiv_rn(acc[t], nelems);
    }
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_correlation_output_shape(int C, int H, int W, int pad_size, int kernel_size, int max_displacement,
                                            int stride1, int stride2, int* out_channels, int* out_h, int* out_w) {
  (void)C;
  if (H <= 0 || W <= 0 || pad_size < 0 || kernel_size < 1 || (kernel_size & 1) == 0 || max_displacement < 0 || stride1 < 1 ||
      stride2 < 1 || !out_channels || !out_h || !out_w)
    return VSR_ERR_INVALID_ARG;
  // correlation_cuda.cc:25-34
  const int border = (kernel_size - 1) / 2 + max_displacement;
  const int ph = H + 2 * pad_size, pw = W + 2 * pad_size;
  const int D = (max_displacement / stride2) * 2 + 1;
  *out_channels = D * D;
  *out_h = (ph - 2 * border + stride1 - 1) / stride1;
  *out_w = (pw - 2 * border + stride1 - 1) / stride1;
  if (*out_h <= 0 || *out_w <= 0) return VSR_ERR_UNSUPPORTED;
  return VSR_OK;
}

extern "C" int vsr_correlation_forward(const float* input1, const float* input2, float* output, int B, int C, int H, int W,
                                       int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                                       int corr_multiply, vsr_stream_t stream) {
  (void)corr_multiply;   // accepted and unused, as in correlation_cuda_kernel.cu:73-147
  if (!input1 || !input2 || !output || B <= 0 || C <= 0) return VSR_ERR_INVALID_ARG;
  int oc, oh, ow;
  int rc = vsr_correlation_output_shape(C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, &oc, &oh, &ow);
  if (rc) return rc;
  const int R = max_displacement / stride2, D = 2 * R + 1;
  if (D > kMaxD) return VSR_ERR_UNSUPPORTED;
  const int segs = ceil_div(ow, 32);
  const int64_t n_tasks = (int64_t)B * oh * D * segs;
  int64_t blocks = ceil_div64(n_tasks, kThreads / 32);
  const int64_t cap = (int64_t)kNumSMs * 8 * 8;
  if (blocks > cap) blocks = cap;
  correlation_forward_kernel<<<(int)blocks, kThreads, 0, as_stream(stream)>>>(input1, input2, output, B, C, H, W, oh, ow,
                                                                             pad_size, kernel_size, max_displacement,
                                                                             stride1, stride2, R, D, segs);
  return after_launch();
}
