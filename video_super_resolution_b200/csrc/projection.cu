// projection.cu -- forward flow projection (a1) and inverse-depth-weighted projection (a2):
// splat with accumulation + count, normalise, hole mask, 4-direction hole fill.
//
// Contract: SURVEY.md Appendix B (the reference ships only the Module surfaces,
// FlowProjectionModule.py:18-33 / DepthProjectionModule.py:12-18; the splat itself has no
// reference implementation -- parity unpinned, oracle/oracle.c::or_flow_projection is the spec).
//
// The four targets of a source pixel all receive the SAME value (Appendix B step 2: no bilinear weights), so the
// splat factors into (1) a histogram over CELLS -- every source adds {-fx*D, -fy*D, D, 1} once to cell
// (int(y2), int(x2)) -- and (2) a 2x2 box sum: target (ty,tx) = cell(ty,tx) + cell(ty,tx-1) + cell(ty-1,tx) +
// cell(ty-1,tx-1), the clamped duplicates of the last row / column being a multiplicity 2.  A cell is one float4;
// the count rides along as a float (exact below 2^24) and is exported as int32 -> count / hole bit-exact.
//
// Two paths:
//  * BOUNDED displacement (caller promises |fx|,|fy| <= max_disp <= 16 px; the pipeline's smooth +-8 px fields):
//    owner-computes in shared memory.  A CTA owns a 192x64 target tile; its 193x65 cells live in shared memory,
//    split into 12 rectangles, one per warp.  A warp scans the sources that can reach its rectangle (rectangle +-
//    bound) and adds the ones that do with plain LDS/STS read-modify-writes -- nobody else touches those cells, so
//    there are no atomics; two lanes of one instruction that hit the same cell are serialised by a one-byte claim
//    protocol.  Box sum + normalise + masks + bitmaps then stream out of shared memory in one coalesced pass.  No
//    global accumulator, no memset, no second pass over the cells: HBM sees the algorithmic bytes only, the L2 the
//    sources ~2.3x.  A source that breaks the promise raises a flag and the whole batch is redone by the general
//    path (one gated launch that exits at once otherwise).
//  * GENERAL (config C3's +-64 px): scatter with one 16-byte red.global.add.v4.f32 per source into an L2-resident
//    cell array, then a gather pass.  The four stages of an image (zero, splat, normalise, fill) are software-
//    pipelined over the batch inside ONE cooperative persistent kernel: in phase p the grid zeroes the cells of image
//    p+1, splats image p, normalises image p-1 and fills image p-2, on three rotating cell arrays (3 x 33 MB at
//    1080p, L2-resident), with a grid barrier between phases.  Round 1 ran four dependent launches per image
//    (memset 6 + splat 13 + normalise 15 + fill 9-19 us at 1080p) and nothing overlapped.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vsr {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 8;  // rows walked by one splat warp task (all of their flow / depth loads are issued up front)

// Reads of data another CTA wrote earlier in the SAME (persistent) kernel go to L2 (ld.global.cg): the non-coherent
// path of __ldg / const __restrict__ may serve a line this SM cached from the previous use of a rotating buffer.
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ float4 shfl_up1_f4(float4 v) {
  return make_float4(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1),
                     __shfl_up_sync(0xffffffffu, v.z, 1), __shfl_up_sync(0xffffffffu, v.w, 1));
}
__device__ __forceinline__ void red_add_f4(float4* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Everything the kernels need to find the buffers of image b.
struct ProjArgs {
  const float* flow;       // (B,h,w,2)
  const float* inv_depth;  // (B,h,w) or nullptr
  float* proj;             // (B,h,w,2)
  float* wsum;             // (B,h,w) or nullptr
  int32_t* count;          // (B,h,w)
  uint8_t* hole;           // (B,h,w)
  float4* acc;             // 3 cell arrays of h*w float4 (general path)
  uint32_t* rowmask;       // per image: h * ceil(w/32) words, bit x%32 of word (y, x/32) = pixel has hits
  uint32_t* colmask;       // per image: ceil(h/32) * w words, bit y%32 of word (y/32, x)
  int* flags;              // [0] promise broken (bounded path), [1 + b] image b has holes
  int B, h, w;
  int bound;               // bounded path: ceil(max_disp)
  int gate;                // general path: 1 = run only if flags[0] != 0 (fallback of the bounded path)
};

__device__ __forceinline__ int64_t rowmask_words(int h, int w) { return (int64_t)h * ceil_div(w, 32); }
__device__ __forceinline__ int64_t colmask_words(int h, int w) { return (int64_t)ceil_div(h, 32) * w; }

// Normalisation of one target from its box sum a = {sum -fx*D, sum -fy*D, sum D, count}.  A target is a hole when
// nothing hit it, and also when the hits carry no usable weight (sum D <= 0 or NaN: an inverse depth of 0 from a
// real estimator) -- dividing by it would put NaN into the warp and the whole conv stack.
__device__ __forceinline__ bool normalise_target(const float4 a, float2& o) {
  const bool is_hole = !(a.w > 0.0f) || !(a.z > 0.0f);
  o = make_float2(0.f, 0.f);
  if (!is_hole) {
    const float rz = __frcp_rn(a.z);     // one reciprocal, two products: <= 1.5 ulp from the quotients
    o = make_float2(a.x * rz, a.y * rz);
  }
  return is_hole;
}

// ---------------------------------------------------------------------------------------------
// GENERAL path, stage roles.  Each role spreads its work over `nw` warps (the whole grid) by a warp-stride loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void role_zero(float4* __restrict__ acc, int64_t n, int64_t tid, int64_t nthreads) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = tid; i < n; i += nthreads) acc[i] = z;
}

// One warp task = 32 consecutive x, kRows consecutive rows.
__device__ __forceinline__ void role_splat(const float* __restrict__ flow, const float* __restrict__ inv_depth,
                                           float4* __restrict__ acc, int h, int w, int warp0, int nw, int lane) {
  const int warps_x = ceil_div(w, 32);
  const int n_tasks = warps_x * ceil_div(h, kRows);
  const float xmax = (float)(w - 1), ymax = (float)(h - 1);
  for (int task = warp0; task < n_tasks; task += nw) {
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

// 2x2 box sum of the cells + normalise + hole mask, and the occupancy bitmaps the fill uses.  One warp task = one
// 32-wide, 16-tall strip: lane = column, the warp walks the rows.  The cell above is the previous row's own cell (a
// register), the cells to the left come from the neighbouring lane by shuffle (lane 0 loads them), so every cell is
// loaded once; 4 rows of loads are in flight at a time.  The ballot of "has hits" is the row word; each lane collects
// its own column bits and writes its 16-bit half of the column word -- no shared memory, no block barrier.  Hole
// pixels get (0,0) here; the fill overwrites them.
constexpr int kStrip = 16, kBatch = 4;

__device__ __forceinline__ void role_normalise(const float4* __restrict__ acc, float* __restrict__ proj,
                                               float* __restrict__ wsum, int32_t* __restrict__ count,
                                               uint8_t* __restrict__ hole, uint32_t* __restrict__ rowmask,
                                               uint32_t* __restrict__ colmask, int* __restrict__ has_holes, int h, int w,
                                               int warp0, int nw, int lane) {
  const int tiles_x = ceil_div(w, 32);
  const int n_strips = 2 * ceil_div(h, 32);          // both halves of every column word get written
  const int n_tasks = tiles_x * n_strips;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int task = warp0; task < n_tasks; task += nw) {
    const int tx = task % tiles_x, strip = task / tiles_x;
    const int x = tx * 32 + lane, y0 = strip * kStrip;
    const bool in_x = x < w;
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
          float2 o;
          is_hole = normalise_target(a, o);
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
      *reinterpret_cast<volatile int*>(has_holes) = 1;   // a flag, not a count: plain store
  }
}

// 4-direction fill (Appendix B step 4): nearest pixel with hits to the left, right, up and down; mean of the found
// (1-4) normalised values, summed in that order; (0,0) if none.  The searches run on the bitmaps, 32 pixels per step.
// One warp per 32-pixel row word: a word without holes costs one 4-byte load for the whole warp; the hole pixels of a
// word are filled in parallel by their lanes.  Only non-hole pixels are read, so the result does not depend on
// execution order.
__device__ __forceinline__ void role_fill(const uint32_t* rowmask_, const uint32_t* colmask_,
                                          float* proj, int h, int w, int warp0, int nw, int lane) {
  struct CG { const uint32_t* p; __device__ __forceinline__ uint32_t operator[](int64_t i) const { return __ldcg(p + i); } };
  const CG rowmask{rowmask_}, colmask{colmask_};
  const int wpr = ceil_div(w, 32), hpr = ceil_div(h, 32);
  const float2* pin = reinterpret_cast<const float2*>(proj);   // non-hole pixels only: never written here
  const int n_words = h * wpr;
  for (int wd = warp0; wd < n_words; wd += nw) {
    const int y = wd / wpr, seg0 = wd - y * wpr;
    const int x = seg0 * 32 + lane;
    const uint32_t hits = rowmask[wd];
    if (x >= w || ((hits >> lane) & 1u)) continue;
    const int p = y * w + x;
    float sx = 0.f, sy = 0.f;
    int found = 0;
    auto take = [&](int yy, int xx) {
      const float2 q = __ldcg(pin + yy * w + xx);   // the neighbour's normalised value, as written by the normalise pass
      sx += q.x;
      sy += q.y;
      ++found;
    };
    {  // left
      int seg = x >> 5;
      uint32_t word = hits & ((1u << (x & 31)) - 1u);
      while (word == 0 && seg > 0) word = rowmask[y * wpr + --seg];
      if (word) take(y, seg * 32 + 31 - __clz(word));
    }
    {  // right
      int seg = x >> 5;
      uint32_t word = hits & ~((2u << (x & 31)) - 1u);
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

// The general path: one cooperative persistent kernel over the whole batch.  Phase p (p = -1 .. B+1):
//   zero the cell array of image p+1 | splat image p | normalise image p-1 | fill image p-2, then a grid barrier.
// Every CTA does its slice of every role; CTAs start at different roles so that at any moment the chip runs a mix of
// the four (the reduction issue rate bounds the splat, L2 reads + HBM writes the normalise, latency the fill).
__global__ void __launch_bounds__(kThreads, 3)
projection_pipeline_kernel(const ProjArgs a) {
  if (a.gate && *reinterpret_cast<volatile int*>(a.flags) == 0) return;   // uniform: nobody reaches a barrier
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * kThreads) >> 5;
  const int warp0 = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nthreads = (int64_t)gridDim.x * kThreads;
  const int64_t P = (int64_t)a.h * a.w;
  const int64_t rw = rowmask_words(a.h, a.w), cw = colmask_words(a.h, a.w);
  for (int p = -1; p <= a.B + 1; ++p) {
    if (p == -1)
      for (int64_t i = tid; i < a.B; i += nthreads) a.flags[1 + i] = 0;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
      const int role = (k + blockIdx.x) & 3;
      if (role == 0) {
        if (p + 1 < a.B) role_zero(a.acc + ((p + 1) % 3) * P, P, tid, nthreads);
      } else if (role == 1) {
        if (p >= 0 && p < a.B)
          role_splat(a.flow + p * P * 2, a.inv_depth ? a.inv_depth + p * P : nullptr, a.acc + (p % 3) * P, a.h, a.w,
                     warp0, nw, lane);
      } else if (role == 2) {
        const int b = p - 1;
        if (b >= 0 && b < a.B)
          role_normalise(a.acc + (b % 3) * P, a.proj + b * P * 2, a.wsum ? a.wsum + b * P : nullptr, a.count + b * P,
                         a.hole + b * P, a.rowmask + b * rw, a.colmask + b * cw, a.flags + 1 + b, a.h, a.w, warp0, nw,
                         lane);
      } else {
        const int b = p - 2;
        if (b >= 0 && b < a.B && *reinterpret_cast<volatile int*>(a.flags + 1 + b) != 0)
          role_fill(a.rowmask + b * rw, a.colmask + b * cw, a.proj + b * P * 2, a.h, a.w, warp0, nw, lane);
      }
    }
    if (p <= a.B) grid.sync();
  }
}

// ---------------------------------------------------------------------------------------------
// BOUNDED path: owner-computes tiles in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kTileThreads = 384;                 // 12 warps
constexpr int kTileW = 192, kTileH = 64;          // targets per tile = 6 x 2 output sub-blocks of 32 x 32
constexpr int kCellW = kTileW + 1, kCellH = kTileH + 1;   // + the column to the left and the row above
constexpr int kCells = kCellW * kCellH;           // 12 545 cells: 200 720 B of float4 + 12 545 B of claim bytes
constexpr int kMaxBound = 16;
constexpr size_t kTileSmem = (size_t)kCells * 16 + ((kCells + 15) / 16) * 16;

__global__ void __launch_bounds__(kTileThreads, 1)
projection_tiled_kernel(const ProjArgs a) {
  extern __shared__ float4 s_cells[];
  uint8_t* s_claim = reinterpret_cast<uint8_t*>(s_cells + kCells);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = a.h, w = a.w, D = a.bound;
  const int tiles_x = ceil_div(w, kTileW), tiles_y = ceil_div(h, kTileH);
  const int n_tiles = tiles_x * tiles_y * a.B;
  const int64_t P = (int64_t)h * w;
  const float xmax = (float)(w - 1), ymax = (float)(h - 1), fD = (float)D;
  const int rw_x = ceil_div(w, 32);
  const int64_t rw = rowmask_words(h, w), cw = colmask_words(h, w);
  // accumulation rectangle of this warp inside the cell tile: 4 x 3 rectangles of 48(49) x 22(21) cells
  const int bx = warp & 3, by = warp >> 2;
  const int cx_lo = 48 * bx, cx_hi = bx == 3 ? kCellW : 48 * (bx + 1);
  const int cy_lo = 22 * by, cy_hi = by == 2 ? kCellH : 22 * (by + 1);
  // output sub-block of this warp: 32 x 32 targets
  const int ox = (warp % 6) * 32, oy = (warp / 6) * 32;
  bool broke = false;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tb = tile / (tiles_x * tiles_y);
    const int tr = tile - tb * tiles_x * tiles_y;
    const int tx0 = (tr % tiles_x) * kTileW, ty0 = (tr / tiles_x) * kTileH;   // first target of the tile
    const float2* flow = reinterpret_cast<const float2*>(a.flow) + tb * P;
    const float* depth = a.inv_depth ? a.inv_depth + tb * P : nullptr;

    for (int i = threadIdx.x; i < kCells; i += kTileThreads) s_cells[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // ---- accumulate: sources of the window [rectangle - D, rectangle + D] (global cell (cy,cx) = local + (ty0-1, tx0-1))
    {
      const int gx_lo = max(tx0 - 1 + cx_lo - D, 0), gx_hi = min(tx0 - 1 + cx_hi - 1 + D, w - 1);   // inclusive
      const int gy_lo = max(ty0 - 1 + cy_lo - D, 0), gy_hi = min(ty0 - 1 + cy_hi - 1 + D, h - 1);
      const int Ws = gx_hi - gx_lo + 1, Hs = gy_hi - gy_lo + 1;
      const int N = (Ws > 0 && Hs > 0) ? Ws * Hs : 0;
      const float inv_ws = Ws > 0 ? 1.0f / (float)Ws : 0.f;
      constexpr int U = 4;
      for (int i0 = 0; i0 < N; i0 += 32 * U) {
        float2 f[U];
        float d[U];
        int sx[U], sy[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int idx = i0 + u * 32 + lane;
          // row-major walk of the window; (idx + 0.5) / Ws is never within rounding of an integer for idx < 2^20
          const int r = (int)(((float)idx + 0.5f) * inv_ws);
          sy[u] = gy_lo + r;
          sx[u] = gx_lo + idx - r * Ws;
          f[u] = make_float2(0.f, 0.f);
          d[u] = 1.0f;
          if (idx < N) {
            const int64_t sp = (int64_t)sy[u] * w + sx[u];
            f[u] = __ldg(flow + sp);
            if (depth) d[u] = __ldg(depth + sp);
          } else {
            sx[u] = -1;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (i0 + u * 32 >= N) break;      // warp-uniform
          const bool valid = sx[u] >= 0;
          const float x2 = __fadd_rn((float)sx[u], f[u].x);
          const float y2 = __fadd_rn((float)sy[u], f[u].y);
          broke |= valid && (fabsf(f[u].x) > fD || fabsf(f[u].y) > fD);
          const bool ok = valid && x2 >= 0.0f && x2 <= xmax && y2 >= 0.0f && y2 <= ymax;
          const int cx = (int)x2 - (tx0 - 1), cy = (int)y2 - (ty0 - 1);
          bool pending = ok && cx >= cx_lo && cx < cx_hi && cy >= cy_lo && cy < cy_hi;
          const int cell = cy * kCellW + cx;
          const float4 v = make_float4(__fmul_rn(-f[u].x, d[u]), __fmul_rn(-f[u].y, d[u]), d[u], 1.0f);
          // the rectangle's cells are this warp's alone: plain read-modify-write.  Lanes of this instruction that hit
          // the same cell take turns: everybody writes its lane number to the cell's claim byte, the survivor goes.
          while (__any_sync(0xffffffffu, pending)) {
            if (pending) s_claim[cell] = (uint8_t)lane;
            __syncwarp();
            const bool win = pending && s_claim[cell] == (uint8_t)lane;
            if (win) {
              float4 c = s_cells[cell];
              c.x += v.x; c.y += v.y; c.z += v.z; c.w += v.w;
              s_cells[cell] = c;
              pending = false;
            }
            __syncwarp();
          }
        }
      }
    }
    __syncthreads();

    // ---- output: 32 x 32 targets per warp, lane = column, rows walked; local cell of target (ty,tx) = (ty+1-ty0, tx+1-tx0)
    {
      const int x = tx0 + ox + lane;
      const bool in_x = x < w;
      const float mx = (x == w - 1) ? 2.0f : 1.0f;
      const int lx = ox + lane + 1;
      float4 up = s_cells[oy * kCellW + lx];
      float4 up_left = s_cells[oy * kCellW + lx - 1];
      uint32_t colbits = 0;
      bool any_hole = false;
      float* proj = a.proj + tb * P * 2;
      float* wsum = a.wsum ? a.wsum + tb * P : nullptr;
      int32_t* count = a.count + tb * P;
      uint8_t* hole = a.hole + tb * P;
      if (ty0 + oy < h && tx0 + ox < w) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int y = ty0 + oy + r;
          const float4 c11 = s_cells[(oy + r + 1) * kCellW + lx];
          float4 c10 = shfl_up1_f4(c11);
          if (lane == 0) c10 = s_cells[(oy + r + 1) * kCellW + lx - 1];
          const float4 c01 = up, c00 = up_left;
          up = c11;
          up_left = c10;
          const bool in_img = in_x && y < h;
          bool is_hole = false;
          if (in_img) {
            const float my = (y == h - 1) ? 2.0f : 1.0f;
            float4 t;
            t.x = (c11.x * mx + c10.x) * my + (c01.x * mx + c00.x);
            t.y = (c11.y * mx + c10.y) * my + (c01.y * mx + c00.y);
            t.z = (c11.z * mx + c10.z) * my + (c01.z * mx + c00.z);
            t.w = (c11.w * mx + c10.w) * my + (c01.w * mx + c00.w);
            const int64_t p = (int64_t)y * w + x;
            float2 o;
            is_hole = normalise_target(t, o);
            reinterpret_cast<float2*>(proj)[p] = o;
            if (wsum) wsum[p] = is_hole ? 0.0f : t.z;
            count[p] = (int32_t)t.w;
            hole[p] = is_hole ? 1 : 0;
          }
          const uint32_t m = __ballot_sync(0xffffffffu, in_img && !is_hole);
          if (lane == 0 && y < h) a.rowmask[tb * rw + (int64_t)y * rw_x + ((tx0 + ox) >> 5)] = m;
          colbits |= (uint32_t)(in_img && !is_hole) << r;
          any_hole |= in_img && is_hole;
        }
        if (in_x) a.colmask[tb * cw + (int64_t)((ty0 + oy) >> 5) * w + x] = colbits;
        if (__any_sync(0xffffffffu, any_hole) && lane == 0) *reinterpret_cast<volatile int*>(a.flags + 1 + tb) = 1;
      }
    }
    __syncthreads();
  }
  if (__any_sync(0xffffffffu, broke) && lane == 0) *reinterpret_cast<volatile int*>(a.flags) = 1;
}

// fill of the bounded path: all images in one launch; nothing to do for an image without holes, and nothing at all
// when the promise was broken (the general path redoes the batch, its own fill included).
__global__ void __launch_bounds__(kThreads)
projection_fill_kernel(const ProjArgs a) {
  if (*reinterpret_cast<volatile int*>(a.flags) != 0) return;
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * kThreads) >> 5;
  const int warp0 = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const int64_t P = (int64_t)a.h * a.w;
  const int64_t rw = rowmask_words(a.h, a.w), cw = colmask_words(a.h, a.w);
  for (int b = 0; b < a.B; ++b)
    if (*reinterpret_cast<volatile int*>(a.flags + 1 + b) != 0)
      role_fill(a.rowmask + b * rw, a.colmask + b * cw, a.proj + b * P * 2, a.h, a.w, warp0, nw, lane);
}

struct ProjWs {
  size_t acc_bytes, flags_off, row_off, col_off, total;
};
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }
inline ProjWs proj_ws(int B, int h, int w) {
  ProjWs s;
  s.acc_bytes = (size_t)h * w * sizeof(float4);
  s.flags_off = 3 * s.acc_bytes;
  s.row_off = s.flags_off + up256((size_t)(B + 1) * 4);
  s.col_off = s.row_off + up256((size_t)B * ceil_div(w, 32) * h * 4);
  s.total = s.col_off + up256((size_t)B * ceil_div(h, 32) * w * 4);
  return s;
}

// co-resident CTAs of the cooperative kernel on the current device
int pipeline_grid() {
  static int per_sm[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = kNumSMs;
  int occ = dev < 64 ? per_sm[dev] : 0;
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, projection_pipeline_kernel, kThreads, 0) != cudaSuccess || occ < 1)
      occ = 1;
    if (dev < 64) per_sm[dev] = occ;
  }
  return sms * occ;
}

int launch_pipeline(const ProjArgs& a, cudaStream_t st) {
  ProjArgs args = a;
  void* kargs[] = {&args};
  int grid = pipeline_grid();
  // no more CTAs than there are warp tasks in the widest role (small images)
  const int64_t tasks = (int64_t)ceil_div(a.w, 32) * a.h;
  const int64_t want = ceil_div64(tasks, kThreads / 32);
  if (want < grid) grid = (int)(want < 1 ? 1 : want);
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(projection_pipeline_kernel), dim3(grid), dim3(kThreads),
                                              kargs, 0, st);
  if (e != cudaSuccess) return cuda_status(e);
  return after_launch();
}

}  // namespace
}  // namespace vsr

using namespace vsr;

extern "C" size_t vsr_flow_projection_workspace_bytes(int B, int h, int w) {
  if (B <= 0 || h <= 0 || w <= 0) return 0;
  return proj_ws(B, h, w).total;
}

extern "C" int vsr_flow_projection_forward_bounded(const float* flow, const float* inv_depth, float* proj, float* wsum,
                                                   int32_t* count, uint8_t* hole, void* workspace, size_t workspace_bytes,
                                                   int B, int h, int w, float max_disp, vsr_stream_t stream) {
  if (!flow || !proj || !count || !hole || !workspace || B <= 0 || h <= 0 || w <= 0) return VSR_ERR_INVALID_ARG;
  if ((int64_t)h * w > (int64_t)1 << 30) return VSR_ERR_UNSUPPORTED;
  if (workspace_bytes < vsr_flow_projection_workspace_bytes(B, h, w)) return VSR_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) % 16 || reinterpret_cast<uintptr_t>(flow) % 8 ||
      reinterpret_cast<uintptr_t>(proj) % 8)
    return VSR_ERR_INVALID_ARG;
  cudaStream_t st = as_stream(stream);
  const ProjWs ws = proj_ws(B, h, w);
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  ProjArgs a;
  a.flow = flow;
  a.inv_depth = inv_depth;
  a.proj = proj;
  a.wsum = wsum;
  a.count = count;
  a.hole = hole;
  a.acc = reinterpret_cast<float4*>(base);
  a.flags = reinterpret_cast<int*>(base + ws.flags_off);
  a.rowmask = reinterpret_cast<uint32_t*>(base + ws.row_off);
  a.colmask = reinterpret_cast<uint32_t*>(base + ws.col_off);
  a.B = B;
  a.h = h;
  a.w = w;
  a.bound = 0;
  a.gate = 0;
  const bool bounded = max_disp >= 0.0f && max_disp <= (float)kMaxBound;    // NaN / negative / large: general path
  if (!bounded) return launch_pipeline(a, st);

  a.bound = (int)ceilf(max_disp);
  static PerDeviceOnce once;
  int dev;
  if (once.needed(&dev)) {
    cudaError_t e = cudaFuncSetAttribute(projection_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem);
    if (e != cudaSuccess) return cuda_status(e);
    once.mark(dev);
  }
  cudaError_t e = cudaMemsetAsync(a.flags, 0, (size_t)(B + 1) * 4, st);
  if (e != cudaSuccess) return cuda_status(e);
  const int n_tiles = ceil_div(w, kTileW) * ceil_div(h, kTileH) * B;
  projection_tiled_kernel<<<n_tiles < kNumSMs ? n_tiles : kNumSMs, kTileThreads, kTileSmem, st>>>(a);
  int rc = after_launch();
  if (rc) return rc;
  const int64_t fill_warps = (int64_t)h * ceil_div(w, 32);
  int64_t fill_blocks = ceil_div64(fill_warps, kThreads / 32);
  if (fill_blocks > kNumSMs * 8) fill_blocks = kNumSMs * 8;
  projection_fill_kernel<<<(int)fill_blocks, kThreads, 0, st>>>(a);
  rc = after_launch();
  if (rc) return rc;
  a.gate = 1;   // the whole batch again through the general path, only if a source broke the promise
  return launch_pipeline(a, st);
}

extern "C" int vsr_flow_projection_forward(const float* flow, const float* inv_depth, float* proj, float* wsum,
                                           int32_t* count, uint8_t* hole, void* workspace, size_t workspace_bytes,
                                           int B, int h, int w, vsr_stream_t stream) {
  return vsr_flow_projection_forward_bounded(flow, inv_depth, proj, wsum, count, hole, workspace, workspace_bytes, B, h,
                                             w, -1.0f, stream);
}
