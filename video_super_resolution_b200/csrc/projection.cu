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

// ---------------------------------------------------------------------------------------------
// Normalise + hole mask, and the occupancy bitmaps the fill uses.  One CTA = one 32x32-pixel tile,
// one warp per row segment: the ballot of "has hits" is the row word (bit x%32 of word (y, x/32));
// the 32 row words of the tile are transposed in shared memory into column words (bit y%32 of
// word (y/32, x)).  Hole pixels get (0,0) here; fill_kernel overwrites them.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
normalise_mask_kernel(const float4* __restrict__ acc, float* __restrict__ proj, float* __restrict__ wsum,
                      int32_t* __restrict__ count, uint8_t* __restrict__ hole, uint32_t* __restrict__ rowmask,
                      uint32_t* __restrict__ colmask, int* __restrict__ n_holes, int h, int w) {
  __shared__ uint32_t rows[32];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int tiles_x = ceil_div(w, 32);
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int x = tx * 32 + lane, y = ty * 32 + wy;
  const bool in_img = (x < w) && (y < h);
  bool is_hole = false;
  if (in_img) {
    const int p = y * w + x;
    const float4 a = acc[p];
    is_hole = !(a.w > 0.0f);
    float2 o = make_float2(0.f, 0.f);
    if (!is_hole) o = make_float2(__fdiv_rn(a.x, a.z), __fdiv_rn(a.y, a.z));
    reinterpret_cast<float2*>(proj)[p] = o;
    if (wsum) wsum[p] = is_hole ? 0.0f : a.z;
    count[p] = (int32_t)a.w;
    hole[p] = is_hole ? 1 : 0;
  }
  const uint32_t m = __ballot_sync(0xffffffffu, in_img && !is_hole);
  const uint32_t hm = __ballot_sync(0xffffffffu, in_img && is_hole);
  if (lane == 0) {
    rows[wy] = m;
    if (y < h) rowmask[y * tiles_x + tx] = m;
    if (hm) *reinterpret_cast<volatile int*>(n_holes) = 1;   // a flag, not a count: plain store (an atomicAdd
                                                             // here serialised 65k warps on one address)
  }
  __syncthreads();
  if (wy == 0 && x < w) {
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) c |= ((rows[k] >> lane) & 1u) << k;
    colmask[ty * w + x] = c;
  }
}

// 4-direction fill (Appendix B step 4): nearest pixel with hits to the left, right, up and down;
// mean of the found (1-4) normalised values, summed in that order; (0,0) if none.  The searches run
// on the bitmaps, 32 pixels per step (the first version walked the accumulator pixel by pixel: a
// 64-px-wide, 540-px-tall hole band cost 150 us per 1080p image; this one 56 us).  One pixel per
// thread on purpose: 4 or 16 pixels per thread serialise the fills of a hole run and measured
// 1.4-1.8x slower.  Only non-hole pixels are read, so the result does not depend on execution order.
__global__ void __launch_bounds__(kThreads)
fill_kernel(const float4* __restrict__ acc, const uint8_t* __restrict__ hole, const uint32_t* __restrict__ rowmask,
            const uint32_t* __restrict__ colmask, const int* __restrict__ n_holes, float* __restrict__ proj, int h, int w) {
  if (*n_holes == 0) return;
  const int n = h * w;
  const int wpr = ceil_div(w, 32), hpr = ceil_div(h, 32);
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    if (!hole[p]) continue;
    const int y = p / w, x = p - y * w;
    float sx = 0.f, sy = 0.f;
    int found = 0;
    auto take = [&](int yy, int xx) {
      const float4 q = acc[yy * w + xx];
      sx += __fdiv_rn(q.x, q.z);
      sy += __fdiv_rn(q.y, q.z);
      ++found;
    };
    {  // left
      int seg = x >> 5;
      uint32_t word = rowmask[y * wpr + seg] & ((1u << (x & 31)) - 1u);
      while (word == 0 && seg > 0) word = rowmask[y * wpr + --seg];
      if (word) take(y, seg * 32 + 31 - __clz(word));
    }
    {  // right
      int seg = x >> 5;
      uint32_t word = rowmask[y * wpr + seg] & ~((2u << (x & 31)) - 1u);
      while (word == 0 && seg + 1 < wpr) word = rowmask[y * wpr + ++seg];
      if (word) take(y, seg * 32 + __ffs(word) - 1);
    }
    {  // up
      int sb = y >> 5;
      uint32_t word = colmask[sb * w + x] & ((1u << (y & 31)) - 1u);
      while (word == 0 && sb > 0) word = colmask[--sb * w + x];
      if (word) take(sb * 32 + 31 - __clz(word), x);
    }
    {  // down
      int sb = y >> 5;
      uint32_t word = colmask[sb * w + x] & ~((2u << (y & 31)) - 1u);
      while (word == 0 && sb + 1 < hpr) word = colmask[++sb * w + x];
      if (word) take(sb * 32 + __ffs(word) - 1, x);
    }
    if (found > 0) reinterpret_cast<float2*>(proj)[p] = make_float2(__fdiv_rn(sx, (float)found), __fdiv_rn(sy, (float)found));
  }
}

}  // namespace
}  // namespace vsr

using namespace vsr;

namespace {
struct ProjWs {
  size_t acc_bytes, ctr_off, row_off, col_off, total;
};
inline ProjWs proj_ws(int h, int w) {
  ProjWs s;
  s.acc_bytes = (size_t)h * w * sizeof(float4);
  s.ctr_off = s.acc_bytes;                                   // zeroed together with the accumulator
  s.row_off = s.ctr_off + 256;
  s.col_off = s.row_off + (((size_t)ceil_div(w, 32) * h * 4 + 255) / 256) * 256;
  s.total = s.col_off + (((size_t)ceil_div(h, 32) * w * 4 + 255) / 256) * 256;
  return s;
}
}  // namespace

extern "C" size_t vsr_flow_projection_workspace_bytes(int B, int h, int w) {
  (void)B;  // one image's accumulator is reused for the whole batch (it stays L2-resident)
  if (h <= 0 || w <= 0) return 0;
  return proj_ws(h, w).total;
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
  const ProjWs ws = proj_ws(h, w);
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  float4* acc = reinterpret_cast<float4*>(base);
  int* n_holes = reinterpret_cast<int*>(base + ws.ctr_off);
  uint32_t* rowmask = reinterpret_cast<uint32_t*>(base + ws.row_off);
  uint32_t* colmask = reinterpret_cast<uint32_t*>(base + ws.col_off);
  const int64_t P = (int64_t)h * w;
  const int n_tasks = ceil_div(w, 32) * ceil_div(h, kRows);
  const int splat_blocks = ceil_div(n_tasks, kThreads / 32);
  const int norm_blocks = ceil_div(w, 32) * ceil_div(h, 32);
  int fill_blocks = (int)ceil_div64(P, kThreads);
  const int cap = kNumSMs * 8 * 4;
  if (fill_blocks > cap) fill_blocks = cap;
  for (int b = 0; b < B; ++b) {
    cudaError_t e = cudaMemsetAsync(acc, 0, ws.acc_bytes + 256, st);
    if (e != cudaSuccess) return cuda_status(e);
    splat_kernel<<<splat_blocks, kThreads, 0, st>>>(flow + b * P * 2, inv_depth ? inv_depth + b * P : nullptr, acc, h,
                                                    w);
    int rc = after_launch();
    if (rc) return rc;
    normalise_mask_kernel<<<norm_blocks, 1024, 0, st>>>(acc, proj + b * P * 2, wsum ? wsum + b * P : nullptr,
                                                        count + b * P, hole + b * P, rowmask, colmask, n_holes, h, w);
    rc = after_launch();
    if (rc) return rc;
    fill_kernel<<<fill_blocks, kThreads, 0, st>>>(acc, hole + b * P, rowmask, colmask, n_holes, proj + b * P * 2, h, w);
    rc = after_launch();
    if (rc) return rc;
  }
  return VSR_OK;
}
