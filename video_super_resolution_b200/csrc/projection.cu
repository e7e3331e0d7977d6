// projection.cu -- forward flow projection (a1) and inverse-depth-weighted projection (a2):
// splat with accumulation + count, normalise, hole mask, 4-direction hole fill.
//
// Contract: SURVEY.md Appendix B (the reference ships only the Module surfaces,
// FlowProjectionModule.py:18-33 / DepthProjectionModule.py:12-18; the splat itself has no
// reference implementation -- parity unpinned, oracle/oracle.c::or_flow_projection is the spec).
//
// Design (B200):
//  * accumulators are ONE float4 per target pixel {sum -fx*D, sum -fy*D, sum D, count}: a target
//    update is a single 16-byte vector reduction (red.global.add.v4.f32, sm_90+) instead of four
//    scalar atomics; the count rides along as a float (exact below 2^24 hits per pixel) and is
//    converted to the int32 the interface exports by the normalise pass -> count/hole bit-exact.
//  * atomics are aggregated before they reach L2: a warp covers 32 consecutive x of a row, the
//    right-hand targets of lane i are handed to lane i+1 by shuffle when they coincide with its
//    left-hand targets (they do wherever the flow is locally smooth), and each thread walks
//    kRows rows carrying its bottom target into the next row's top target.  A smooth field costs
//    ~1.25 vector reductions per source pixel instead of 16 scalar atomics; a pathological field
//    (config C3, +-64 px i.i.d.) degrades gracefully to 4 vector reductions.
//  * images are processed one at a time with a single-image accumulator (16 B/pixel, 33 MB at
//    1080p) that stays resident in the 126 MB L2, so HBM sees only the algorithmic traffic
//    (flow/depth in, proj/wsum/count/hole out).
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 8;  // rows walked by one thread (all of their flow / depth loads are issued up front)

__device__ __forceinline__ void red_add_f4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 shfl_up_f4(float4 v) {
  return make_float4(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1),
                     __shfl_up_sync(0xffffffffu, v.z, 1), __shfl_up_sync(0xffffffffu, v.w, 1));
}

// One warp = 32 consecutive x, kRows consecutive rows of image `b`.
__global__ void __launch_bounds__(kThreads)
splat_kernel(const float* __restrict__ flow, const float* __restrict__ inv_depth, float4* __restrict__ acc,
             int h, int w) {
  const int lane = threadIdx.x & 31;
  const int warps_x = ceil_div(w, 32);
  const int n_tasks = warps_x * ceil_div(h, kRows);
  const int warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * kThreads) >> 5;

  for (int task = warp_global; task < n_tasks; task += n_warps) {
    const int x = (task % warps_x) * 32 + lane;
    const int y0 = (task / warps_x) * kRows;
    bool carry_valid = false;
    int carry_t = 0;
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);

    // ncu: 48 % of this kernel's stall samples sat on the first use of the flow load -- a warp
    // walked its rows with one dependent load per row.  Issue every row's loads first.
    float2 fl[kRows];
    float dp[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int y = y0 + r;
      fl[r] = make_float2(0.f, 0.f);
      dp[r] = 1.0f;
      if (x < w && y < h) {
        const int64_t p = (int64_t)y * w + x;
        fl[r] = ldg_stream_f2(reinterpret_cast<const float2*>(flow) + p);
        if (inv_depth) dp[r] = __ldg(inv_depth + p);
      }
    }

#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int y = y0 + r;
      bool valid = false;
      int xL = 0, xR = 0, yT = 0, yB = 0;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (x < w && y < h) {
        const float2 f = fl[r];
        float x2 = __fadd_rn((float)x, f.x);
        float y2 = __fadd_rn((float)y, f.y);
        // Appendix B step 2 (the comparison form also rejects NaN)
        if (x2 >= 0.0f && x2 <= (float)(w - 1) && y2 >= 0.0f && y2 <= (float)(h - 1)) {
          float d = dp[r];
          valid = true;
          xL = (int)x2;
          yT = (int)y2;
          xR = min(xL + 1, w - 1);
          yB = min(yT + 1, h - 1);
          v = make_float4(__fmul_rn(-f.x, d), __fmul_rn(-f.y, d), d, 1.0f);
        }
      }
      const bool dupx = (xR == xL), dupy = (yB == yT);  // clamped duplicate targets are hit twice
      const float mT = dupy ? 2.0f : 1.0f;
      float4 topL = f4_scale(v, (dupx ? 2.0f : 1.0f) * mT);
      float4 botL = f4_scale(v, dupx ? 2.0f : 1.0f);
      const bool has_bot = valid && !dupy;
      bool has_R = valid && !dupx;

      // horizontal hand-over: lane i's right column -> lane i+1's left column
      float4 pv = shfl_up_f4(v);
      int p_xR = __shfl_up_sync(0xffffffffu, xR, 1);
      int p_yT = __shfl_up_sync(0xffffffffu, yT, 1);
      int p_yB = __shfl_up_sync(0xffffffffu, yB, 1);
      int p_hasR = __shfl_up_sync(0xffffffffu, (int)has_R, 1);
      bool absorb = valid && lane > 0 && p_hasR && p_xR == xL && p_yT == yT && p_yB == yB;
      if (absorb) {
        topL = f4_add(topL, f4_scale(pv, mT));
        botL = f4_add(botL, pv);
      }
      int absorbed_by_next = __shfl_down_sync(0xffffffffu, (int)absorb, 1);
      if (lane < 31 && absorbed_by_next) has_R = false;

      // vertical carry: previous row's bottom-left target -> this row's top-left target
      if (carry_valid) {
        if (valid && carry_t == yT * w + xL) topL = f4_add(topL, carry);
        else red_add_f4(acc + carry_t, carry);
        carry_valid = false;
      }
      if (valid) {
        red_add_f4(acc + (yT * w + xL), topL);
        if (has_R) {
          red_add_f4(acc + (yT * w + xR), f4_scale(v, mT));
          if (!dupy) red_add_f4(acc + (yB * w + xR), v);
        }
        if (has_bot) {
          carry_valid = true;
          carry_t = yB * w + xL;
          carry = botL;
        }
      }
    }
    if (carry_valid) red_add_f4(acc + carry_t, carry);
  }
}

// 4-direction fill of one hole pixel (Appendix B step 4): nearest non-hole pixel to the left,
// right, up and down; mean of the found (1-4) normalised values in that order; (0,0) if none.
// Only non-hole pixels are read, so the result does not depend on execution order.
__device__ __forceinline__ float2 fill_hole(const float4* __restrict__ acc, int x, int y, int h, int w) {
  float sx = 0.f, sy = 0.f;
  int found = 0;
  for (int xx = x - 1; xx >= 0; --xx) {
    const float4 q = acc[y * w + xx];
    if (q.w > 0.0f) { sx += __fdiv_rn(q.x, q.z); sy += __fdiv_rn(q.y, q.z); ++found; break; }
  }
  for (int xx = x + 1; xx < w; ++xx) {
    const float4 q = acc[y * w + xx];
    if (q.w > 0.0f) { sx += __fdiv_rn(q.x, q.z); sy += __fdiv_rn(q.y, q.z); ++found; break; }
  }
  for (int yy = y - 1; yy >= 0; --yy) {
    const float4 q = acc[yy * w + x];
    if (q.w > 0.0f) { sx += __fdiv_rn(q.x, q.z); sy += __fdiv_rn(q.y, q.z); ++found; break; }
  }
  for (int yy = y + 1; yy < h; ++yy) {
    const float4 q = acc[yy * w + x];
    if (q.w > 0.0f) { sx += __fdiv_rn(q.x, q.z); sy += __fdiv_rn(q.y, q.z); ++found; break; }
  }
  if (found == 0) return make_float2(0.f, 0.f);
  return make_float2(__fdiv_rn(sx, (float)found), __fdiv_rn(sy, (float)found));
}

// Normalise + hole mask + fill for one image; kPix consecutive pixels per thread so that the four
// accumulator loads are in flight together and every output leaves as one vector store
// (kPix = 4: proj 2 x 16 B, wsum 16 B, count 16 B, hole 4 B).  kPix = 1 is the ragged fallback.
template <int kPix>
__global__ void __launch_bounds__(kThreads)
normalise_fill_kernel(const float4* __restrict__ acc, float* __restrict__ proj, float* __restrict__ wsum,
                      int32_t* __restrict__ count, uint8_t* __restrict__ hole, int h, int w) {
  const int n = h * w / kPix;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p0 = i * kPix;
    float4 a[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) a[k] = acc[p0 + k];
    float2 o[kPix];
    float ws[kPix];
    int32_t cn[kPix];
    uint8_t hl[kPix];
#pragma unroll
    for (int k = 0; k < kPix; ++k) {
      const bool is_hole = !(a[k].w > 0.0f);
      if (!is_hole) {
        o[k] = make_float2(__fdiv_rn(a[k].x, a[k].z), __fdiv_rn(a[k].y, a[k].z));
      } else {
        const int p = p0 + k;
        const int y = p / w;
        o[k] = fill_hole(acc, p - y * w, y, h, w);
      }
      ws[k] = is_hole ? 0.0f : a[k].z;
      cn[k] = (int32_t)a[k].w;
      hl[k] = is_hole ? 1 : 0;
    }
    if (kPix == 4) {
      reinterpret_cast<float4*>(proj)[2 * i] = make_float4(o[0].x, o[0].y, o[1 % kPix].x, o[1 % kPix].y);
      reinterpret_cast<float4*>(proj)[2 * i + 1] = make_float4(o[2 % kPix].x, o[2 % kPix].y, o[3 % kPix].x, o[3 % kPix].y);
      if (wsum) reinterpret_cast<float4*>(wsum)[i] = make_float4(ws[0], ws[1 % kPix], ws[2 % kPix], ws[3 % kPix]);
      reinterpret_cast<int4*>(count)[i] = make_int4(cn[0], cn[1 % kPix], cn[2 % kPix], cn[3 % kPix]);
      reinterpret_cast<uchar4*>(hole)[i] = make_uchar4(hl[0], hl[1 % kPix], hl[2 % kPix], hl[3 % kPix]);
    } else {
      reinterpret_cast<float2*>(proj)[p0] = o[0];
      if (wsum) wsum[p0] = ws[0];
      count[p0] = cn[0];
      hole[p0] = hl[0];
    }
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" size_t vsr_flow_projection_workspace_bytes(int B, int h, int w) {
  (void)B;  // one image's accumulator is reused for the whole batch (it stays L2-resident)
  if (h <= 0 || w <= 0) return 0;
  return (size_t)h * (size_t)w * sizeof(float4);
}

extern "C" int vsr_flow_projection_forward(const float* flow, const float* inv_depth, float* proj, float* wsum,
                                           int32_t* count, uint8_t* hole, void* workspace, size_t workspace_bytes,
                                           int B, int h, int w, vsr_stream_t stream) {
  if (!flow || !proj || !count || !hole || !workspace || B <= 0 || h <= 0 || w <= 0) return VSR_ERR_INVALID_ARG;
  if ((int64_t)h * w > (int64_t)1 << 30) return VSR_ERR_UNSUPPORTED;
  if (workspace_bytes < vsr_flow_projection_workspace_bytes(B, h, w)) return VSR_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 16 || reinterpret_cast<uintptr_t>(flow) % 8 ||
      reinterpret_cast<uintptr_t>(proj) % 8)
    return VSR_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  float4* acc = reinterpret_cast<float4*>(workspace);
  const int64_t P = (int64_t)h * w;
  const int n_tasks = ceil_div(w, 32) * ceil_div(h, kRows);
  int splat_blocks = ceil_div(n_tasks, kThreads / 32);
  const uintptr_t al = reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(count) |
                       (wsum ? reinterpret_cast<uintptr_t>(wsum) : 0);
  // measured on B200: 4 pixels per thread is SLOWER (52 -> 70 us per 1080p image; 3x slower on hole-heavy
  // scenes, where the serial fill loops of one thread add up) -- kept for reference, not used.
  const bool vec4 = false && (P % 4 == 0) && (al % 16 == 0) && (reinterpret_cast<uintptr_t>(hole) % 4 == 0);
  int norm_blocks = (int)ceil_div64(vec4 ? P / 4 : P, kThreads);
  const int cap = kNumSMs * 8 * 4;
  if (norm_blocks > cap) norm_blocks = cap;
  for (int b = 0; b < B; ++b) {
    cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)P * sizeof(float4), st);
    if (e != cudaSuccess) return cuda_status(e);
    splat_kernel<<<splat_blocks, kThreads, 0, st>>>(flow + b * P * 2, inv_depth ? inv_depth + b * P : nullptr, acc, h,
                                                    w);
    int rc = after_launch();
    if (rc) return rc;
    if (vec4)
      normalise_fill_kernel<4><<<norm_blocks, kThreads, 0, st>>>(acc, proj + b * P * 2, wsum ? wsum + b * P : nullptr,
                                                                 count + b * P, hole + b * P, h, w);
    else
      normalise_fill_kernel<1><<<norm_blocks, kThreads, 0, st>>>(acc, proj + b * P * 2, wsum ? wsum + b * P : nullptr,
                                                                 count + b * P, hole + b * P, h, w);
    rc = after_launch();
    if (rc) return rc;
  }
  return VSR_OK;
}
