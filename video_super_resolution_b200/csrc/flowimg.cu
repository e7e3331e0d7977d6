// flowimg.cu -- flow colour coding and its resize glue on the device.
//   ref: utils/flow_utils.py:4-24 (flow2img), :27-61 (compute_color), :64-112 (make_color_wheel),
//        my_packages/FlowProjection/FlowProjectionModule.py:31-32 (flow -> .cpu().numpy() -> flow2img -> .cuda()),
//        network/video_super_resolution.py:35,52 (transpose1323 + F.interpolate default nearest)
// The reference colour-codes FlowNet2's output on the host: a D2H copy, ~20 numpy passes and an H2D copy per
// flow map, four times per frame.  Here it is a global max reduction and one elementwise kernel that writes the
// u8 image and/or the transposed, nearest-resized fp32 planes the stack consumes.
//
// Arithmetic: the reference's provider is NumPy; pinned to NumPy >= 2 promotion rules by tests/golden/flow2img.npz
// (rad / maxrad / the normalising division in fp32, everything after `+ np.finfo(float).eps` in float64).  The
// rounding sequence is spelled with _rn intrinsics so that nvcc does not contract it into FMAs; atan2 in double
// differs from the host libm by <= 2 ulp, i.e. a different u8 only if 255*col lands within ~1e-13 of an integer.
#include "common.cuh"

#include <math.h>

namespace vsr {
namespace {

constexpr int kThreads = 256;
__constant__ double c_wheel[55 * 3];   // Middlebury colour wheel / 255 (flow_utils.py:64-112, :51-52)

void make_wheel(double* wheel) {
  const int RY = 15, YG = 6, GC = 4, CB = 11, BM = 13, MR = 6;
  double t[55][3] = {};
  int col = 0;
  for (int i = 0; i < RY; ++i) { t[col + i][0] = 255; t[col + i][1] = floor(255.0 * i / RY); }
  col += RY;
  for (int i = 0; i < YG; ++i) { t[col + i][0] = 255 - floor(255.0 * i / YG); t[col + i][1] = 255; }
  col += YG;
  for (int i = 0; i < GC; ++i) { t[col + i][1] = 255; t[col + i][2] = floor(255.0 * i / GC); }
  col += GC;
  for (int i = 0; i < CB; ++i) { t[col + i][1] = 255 - floor(255.0 * i / CB); t[col + i][2] = 255; }
  col += CB;
  for (int i = 0; i < BM; ++i) { t[col + i][2] = 255; t[col + i][0] = floor(255.0 * i / BM); }
  col += BM;
  for (int i = 0; i < MR; ++i) { t[col + i][2] = 255 - floor(255.0 * i / MR); t[col + i][0] = 255; }
  for (int i = 0; i < 55; ++i)
    for (int c = 0; c < 3; ++c) wheel[i * 3 + c] = t[i][c] / 255;
}

__device__ __forceinline__ bool unknown_flow(float u, float v) { return fabsf(u) > 1e7f || fabsf(v) > 1e7f; }

// ws[0] = bits of max rad (rad >= 0, so the unsigned order of the bits is the float order), ws[1] = NaN seen
__global__ void __launch_bounds__(kThreads)
flow_maxrad_kernel(const float2* __restrict__ flow, int64_t n, unsigned* __restrict__ ws) {
  float m = 0.0f;
  bool nan = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 f = ldg_stream_f2(flow + i);
    if (unknown_flow(f.x, f.y)) f.x = f.y = 0.0f;
    const float r = __fsqrt_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)));
    if (r != r) nan = true;
    else m = fmaxf(m, r);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  nan = __any_sync(0xffffffffu, nan);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(ws, __float_as_uint(m));
    if (nan) atomicOr(ws + 1, 1u);
  }
}

// one thread per OUTPUT pixel (Y,X) of an (out_h,out_w) raster; source pixel by ATen's nearest rule
// (identity when the sizes agree).  img: (h,w,3) u8 at the source raster (only when sizes agree);
// planes: (3,out_h,out_w) fp32.
__global__ void __launch_bounds__(kThreads)
flow_colour_kernel(const float2* __restrict__ flow, const unsigned* __restrict__ ws, uint8_t* __restrict__ img,
                   float* __restrict__ planes, int h, int w, int out_h, int out_w) {
  const float maxrad = ws[1] ? -1.0f : __uint_as_float(ws[0]);   // max(-1, np.max(rad)); Python max(-1, nan) = -1
  const float sy = (float)h / (float)out_h, sx = (float)w / (float)out_w;
  const int64_t n = (int64_t)out_h * out_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int Y = (int)(i / out_w), X = (int)(i - (int64_t)Y * out_w);
    const int y = min((int)floorf(__fmul_rn((float)Y, sy)), h - 1), x = min((int)floorf(__fmul_rn((float)X, sx)), w - 1);
    float2 f = __ldg(flow + (int64_t)y * w + x);
    const bool unknown = unknown_flow(f.x, f.y);
    if (unknown) f.x = f.y = 0.0f;
    double u = __dadd_rn((double)__fdiv_rn(f.x, maxrad), 2.220446049250313e-16);
    double v = __dadd_rn((double)__fdiv_rn(f.y, maxrad), 2.220446049250313e-16);
    const bool isnan_ = (u != u) || (v != v);
    if (isnan_) u = v = 0.0;
    const double rad = __dsqrt_rn(__dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v)));
    const double a = __ddiv_rn(atan2(-v, -u), 3.141592653589793);
    const double fk = __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(a, 1.0), 2.0), 54.0), 1.0);
    const double k0f = floor(fk);
    const int k0 = (int)k0f;
    int k1 = k0 + 1;
    if (k1 == 56) k1 = 1;
    const double fr = __dadd_rn(fk, -k0f);
    uint8_t px[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double col0 = c_wheel[(k0 - 1) * 3 + c], col1 = c_wheel[(k1 - 1) * 3 + c];
      double col = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -fr), col0), __dmul_rn(fr, col1));
      if (rad <= 1.0) col = __dadd_rn(1.0, -__dmul_rn(rad, __dadd_rn(1.0, -col)));
      else col = __dmul_rn(col, 0.75);
      const double val = floor(__dmul_rn(__dmul_rn(255.0, col), isnan_ ? 0.0 : 1.0));
      px[c] = unknown ? (uint8_t)0 : (uint8_t)(int)val;
    }
    if (img) {
      img[i * 3 + 0] = px[0];
      img[i * 3 + 1] = px[1];
      img[i * 3 + 2] = px[2];
    }
    if (planes) {
      planes[i] = (float)px[0];
      planes[n + i] = (float)px[1];
      planes[2 * n + i] = (float)px[2];
    }
  }
}

inline int grid_for(int64_t n) {
  int64_t blocks = ceil_div64(n, kThreads);
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" int vsr_flow_to_image(const float* flow, int h, int w, uint8_t* img_u8, float* planes, int out_h, int out_w,
                                 void* workspace, vsr_stream_t stream) {
  if (!flow || !workspace || h <= 0 || w <= 0 || (!img_u8 && !planes)) return VSR_ERR_INVALID_ARG;
  if (planes && (out_h <= 0 || out_w <= 0)) return VSR_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  static PerDeviceOnce once;   // __constant__ data is per device; 1.3 KB
  int dev;
  if (once.needed(&dev)) {
    double wheel[55 * 3];
    make_wheel(wheel);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_wheel, wheel, sizeof(wheel), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_status(e);
    e = cudaStreamSynchronize(st);   // `wheel` lives on this stack frame
    if (e != cudaSuccess) return cuda_status(e);
    once.mark(dev);
  }
  cudaError_t e = cudaMemsetAsync(workspace, 0, 8, st);
  if (e != cudaSuccess) return cuda_status(e);
  const int64_t n = (int64_t)h * w;
  unsigned* ws = reinterpret_cast<unsigned*>(workspace);
  flow_maxrad_kernel<<<grid_for(n), kThreads, 0, st>>>(reinterpret_cast<const float2*>(flow), n, ws);
  int rc = after_launch();
  if (rc) return rc;
  const bool same = planes && out_h == h && out_w == w;
  if (img_u8 || same) {
    flow_colour_kernel<<<grid_for(n), kThreads, 0, st>>>(reinterpret_cast<const float2*>(flow), ws, img_u8,
                                                         same ? planes : nullptr, h, w, h, w);
    rc = after_launch();
    if (rc) return rc;
  }
  if (planes && !same) {
    flow_colour_kernel<<<grid_for((int64_t)out_h * out_w), kThreads, 0, st>>>(
        reinterpret_cast<const float2*>(flow), ws, nullptr, planes, h, w, out_h, out_w);
    rc = after_launch();
  }
  return rc;
}
