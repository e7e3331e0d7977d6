// projection.cu -- forward flow projection (a1) and inverse-depth-weighted projection (a2):
// splat with accumulation + count, normalise, hole mask, 4-direction hole fill.
//
// Contract: SURVEY.md Appendix B (the reference ships only the Module surfaces,
// FlowProjectionModule.py:18-33 / DepthProjectionModule.py:12-18; the splat itself has no
// reference implementation -- parity unpinned, oracle/oracle.c::or_flow_projection is the spec).
//
// Design (B200):
//  * The four targets of a source pixel all receive the SAME value (Appendix B step 2: no bilinear
//    weights), so the splat factors into (1) a histogram over CELLS -- every source adds its value once to
//    cell (yT, xL) = (int(y2), int(x2)) -- and (2) a 2x2 box sum: target (ty,tx) = cell(ty,tx) + cell(ty,tx-1)
//    + cell(ty-1,tx) + cell(ty-1,tx-1), with the clamped duplicates (xR == xL at the last column, yB == yT at
//    the last row) as multiplicity 2 of the target's own column / row.  One vector reduction per source
//    pixel instead of four (the first version aggregated neighbouring lanes' targets by shuffle and still
//    paid ~1.25 on smooth fields and 4 on config C3's i.i.d. field); the box sum is a gather (registers +
//    one shuffle) fused into the normalise pass, in a fixed order.
//  * A cell is ONE float4 {sum -fx*D, sum -fy*D, sum D, count}: a 16-byte red.global.add.v4.f32 (sm_90+);
//    the count rides along as a float (exact below 2^24) and is exported as int32 -> count/hole bit-exact.
//  * Images are processed one at a time with a single-image cell array (16 B/pixel, 33 MB at 1080p) that
//    stays resident in the 126 MB L2, so HBM sees only the algorithmic traffic (flow/depth in,
//    proj/wsum/count/hole out).
#include "common.cuh"

namespace vsr {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 8;  // rows walked by one thread (all of their flow / depth loads are issued up front)

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 shfl_up1_f4(float4 v) {
  return make_float4(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1),
                     __shfl_up_sync(0xffffffffu, v.z, 1), __shfl_up_sync(0xffffffffu, v.w, 1));
}
__device__ __forceinline__ void red_add_f4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// One warp = 32 consecutive x, kRows consecutive rows of one image.
__global__ void __launch_bounds__(kThreads)
splat_kernel(const float* __restrict__ flow, const float* __restrict__ inv_depth, float4* __restrict__ acc,
             int h, int w) {
  const int lane = threadIdx.x & 31;
  const int warps_x = ceil_div(w, 32);
  const int n_tasks = warps_x * ceil_div(h, kRows);
  const int warp_global = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int n_warps = (gridDim.x * kThreads) >> 5;
  const float xmax = (float)(w - 1), ymax = (float)(h - 1);

  for (int task = warp_global; task < n_tasks; task += n_warps) {
    const int x = (task % warps_x) * 32 + lane;
    const int y0 = (task / warps_x) * kRows;
    // ncu (first version): 48 % of the stall samples sat on the first use of the flow load -- a warp
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
      if (x < w && y < h) {
        const float2 f = fl[r];
        const float x2 = __fadd_rn((float)x, f.x);
        const float y2 = __fadd_rn((float)y, f.y);
        // Appendix B step 2 (the comparison form also rejects NaN)
        if (x2 >= 0.0f && x2 <= xmax && y2 >= 0.0f && y2 <= ymax) {
          const float d = dp[r];
          red_add_f4(acc + ((int)y2 * w + (int)x2), make_float4(__fmul_rn(-f.x, d), __fmul_rn(-f.y, d), d, 1.0f));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 2x2 box sum of the cells + normalise + hole mask, and the occupancy bitmaps the fill uses.
// One WARP = one 32-wide, 16-tall strip: lane = column, the warp walks the rows.  The cell above is the
// previous row's own cell (a register), the cells to the left come from the neighbouring lane by shuffle
// (lane 0 loads them), so every cell is loaded once; 8 rows of loads are in flight at a time.  The ballot of
// "has hits" is the row word (bit x%32 of word (y, x/32)); each lane collects its own column bits (bit y%32
// of word (y/32, x)) and writes its 16-bit half of the column word -- no shared memory, no block barrier
// (the first version used 32x32 tiles of 1024 threads with a transposition through shared memory and spent
// 20 us per 1080p image in load -> barrier -> store waves).  Hole pixels get (0,0) here; fill_kernel
// overwrites them.
// ---------------------------------------------------------------------------------------------
constexpr int kStrip = 16, kBatch = 4;   // 4 rows of loads in flight: 8 needed 98 registers (2 CTAs per SM)

__global__ void __launch_bounds__(kThreads, 4)
normalise_mask_kernel(const float4* __restrict__ acc, float* __restrict__ proj, float* __restrict__ wsum,
                      int32_t* __restrict__ count, uint8_t* __restrict__ hole, uint32_t* __restrict__ rowmask,
                      uint32_t* __restrict__ colmask, int* __restrict__ n_holes, int h, int w) {
  const int lane = threadIdx.x & 31;
  const int tiles_x = ceil_div(w, 32);
  const int n_strips = 2 * ceil_div(h, 32);          // both halves of every column word get written
  const int task = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  if (task >= tiles_x * n_strips) return;
  const int tx = task % tiles_x, strip = task / tiles_x;
  const int x = tx * 32 + lane, y0 = strip * kStrip;
  const bool in_x = x < w;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float mx = (x == w - 1) ? 2.0f : 1.0f;
  float4 up = (in_x && y0 > 0 && y0 <= h) ? ldg_f4(acc + (y0 - 1) * w + x) : zero4;            // cell (y0-1, x)
  float4 up_left = shfl_up1_f4(up);                                                              // cell (y0-1, x-1)
  if (lane == 0) up_left = (x > 0 && y0 > 0 && y0 <= h) ? ldg_f4(acc + (y0 - 1) * w + x - 1) : zero4;
  uint32_t colbits = 0;
  bool any_hole = false;
#pragma unroll
  for (int r0 = 0; r0 < kStrip; r0 += kBatch) {
    float4 c[kBatch], cl[kBatch];
#pragma unroll
    for (int r = 0; r < kBatch; ++r) {
      const int y = y0 + r0 + r;
      c[r] = (in_x && y < h) ? ldg_f4(acc + y * w + x) : zero4;
      cl[r] = (lane == 0 && x > 0 && y < h) ? ldg_f4(acc + y * w + x - 1) : zero4;
    }
#pragma unroll
    for (int r = 0; r < kBatch; ++r) {
      const int y = y0 + r0 + r;
      const float4 c11 = c[r], c01 = up;
      float4 c10 = shfl_up1_f4(c11);
      if (lane == 0) c10 = cl[r];
      const float4 c00 = up_left;            // cell (y-1, x-1) = the previous row's left neighbour
      up = c11;
      up_left = c10;
      const bool in_img = in_x && y < h;
      bool is_hole = false;
      if (in_img) {
        // 2x2 box sum in a fixed order; the clamped duplicate targets (Appendix B: "hit twice") are the
        // multiplicity 2 of the target's own column at x = w-1 and of its own row at y = h-1
        const float my = (y == h - 1) ? 2.0f : 1.0f;
        float4 a;
        a.x = (c11.x * mx + c10.x) * my + (c01.x * mx + c00.x);
        a.y = (c11.y * mx + c10.y) * my + (c01.y * mx + c00.y);
        a.z = (c11.z * mx + c10.z) * my + (c01.z * mx + c00.z);
        a.w = (c11.w * mx + c10.w) * my + (c01.w * mx + c00.w);
        const int p = y * w + x;
        is_hole = !(a.w > 0.0f);
        float2 o = make_float2(0.f, 0.f);
        if (!is_hole) {
          const float rz = __frcp_rn(a.z);     // one reciprocal, two products: <= 1.5 ulp from the quotients
          o = make_float2(a.x * rz, a.y * rz);
        }
        reinterpret_cast<float2*>(proj)[p] = o;
        if (wsum) wsum[p] = is_hole ? 0.0f : a.z;
        count[p] = (int32_t)a.w;
        hole[p] = is_hole ? 1 : 0;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, in_img && !is_hole);
      if (lane == 0 && y < h) rowmask[y * tiles_x + tx] = m;
      colbits |= (uint32_t)(in_img && !is_hole) << (r0 + r);
      any_hole |= in_img && is_hole;
    }
  }
  if (in_x) reinterpret_cast<uint16_t*>(colmask)[((strip >> 1) * w + x) * 2 + (strip & 1)] = (uint16_t)colbits;
  if (__any_sync(0xffffffffu, any_hole) && lane == 0)
    *reinterpret_cast<volatile int*>(n_holes) = 1;   // a flag, not a count: plain store (an atomicAdd here
                                                     // serialised 65k warps on one address)
}

// 4-direction fill (Appendix B step 4): nearest pixel with hits to the left, right, up and down;
// mean of the found (1-4) normalised values, summed in that order; (0,0) if none.  The searches run
// on the bitmaps, 32 pixels per step (the first version walked the accumulator pixel by pixel: a
// 64-px-wide, 540-px-tall hole band cost 150 us per 1080p image; this one 56 us).  One pixel per
// thread on purpose: 4 or 16 pixels per thread serialise the fills of a hole run and measured
// 1.4-1.8x slower.  Only non-hole pixels are read, so the result does not depend on execution order.
__global__ void __launch_bounds__(kThreads)
fill_kernel(const uint32_t* __restrict__ rowmask, const uint32_t* __restrict__ colmask, const int* __restrict__ n_holes,
            float* proj, int h, int w) {
  if (*n_holes == 0) return;
  const int wpr = ceil_div(w, 32), hpr = ceil_div(h, 32);
  const float2* pin = reinterpret_cast<const float2*>(proj);   // non-hole pixels only: never written here
  const int lane = threadIdx.x & 31;
  const int n_words = h * wpr;
  // one warp per 32-pixel row word (the grid gives every warp exactly one word, so there is no chain of dependent
  // loads): a word without holes costs one 4-byte load for the whole warp; the hole pixels of a word are filled
  // in parallel by their lanes.  (Visiting 32 words per warp and only those with holes serialises the searches
  // of scattered holes: 46 us instead of 9 us on a smooth field.)
  for (int wd = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; wd < n_words; wd += (gridDim.x * blockDim.x) >> 5) {
    {
      const int y = wd / wpr, seg0 = wd - y * wpr;
      const int x = seg0 * 32 + lane;
      const uint32_t hits = rowmask[wd];
      if (x >= w || ((hits >> lane) & 1u)) continue;
      const int p = y * w + x;
    float sx = 0.f, sy = 0.f;
    int found = 0;
    auto take = [&](int yy, int xx) {
      const float2 q = pin[yy * w + xx];            // the neighbour's normalised value, as written by the normalise pass
      sx += q.x;
      sy += q.y;
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
  const int norm_blocks = ceil_div(ceil_div(w, 32) * 2 * ceil_div(h, 32), kThreads / 32);
  const int fill_blocks = ceil_div(h * ceil_div(w, 32), kThreads / 32);   // one warp per 32-pixel row word
  for (int b = 0; b < B; ++b) {
    cudaError_t e = cudaMemsetAsync(acc, 0, ws.acc_bytes + 256, st);   // cells + hole flag
    if (e != cudaSuccess) return cuda_status(e);
    splat_kernel<<<splat_blocks, kThreads, 0, st>>>(flow + b * P * 2, inv_depth ? inv_depth + b * P : nullptr, acc, h,
                                                    w);
    int rc = after_launch();
    if (rc) return rc;
    normalise_mask_kernel<<<norm_blocks, kThreads, 0, st>>>(acc, proj + b * P * 2, wsum ? wsum + b * P : nullptr,
                                                        count + b * P, hole + b * P, rowmask, colmask, n_holes, h, w);
    rc = after_launch();
    if (rc) return rc;
    fill_kernel<<<fill_blocks, kThreads, 0, st>>>(rowmask, colmask, n_holes, proj + b * P * 2, h, w);
    rc = after_launch();
    if (rc) return rc;
  }
  return VSR_OK;
}
