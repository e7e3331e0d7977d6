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
//  * GENERAL (config C3's +-64 px): scatter with one 16-byte red.global.add.v4.f32 per source into ONE L2-resident
//    cell array (33 MB at 1080p), then a gather pass (box sum + normalise + masks) and the fill, image after image on
//    the caller's stream.  Overlapping the stages of successive images was tried three ways and measured slower (see
//    run_general); the cooperative pipelined kernel of those experiments survives as the gated fallback of the
//    bounded path, where what matters is that skipping it costs one empty launch.
#include <cooperative_groups.h>


#include "common.cuh"

namespace cg = cooperative_groups;

namespace vsr {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 8;  // rows walked by one splat warp task (all of their flow / depth loads are issued up front)

// Reads of data another CTA wrote earlier in the SAME (persistent) kernel go to L2 (ld.global.cg): the non-coherent
// path of __ldg / const __restrict__ may serve a line this SM cached from the previous use of a rotating buffer.
template <bool CG> __device__ __forceinline__ float4 ldg_f4(const float4* p) { return CG ? __ldcg(p) : __ldg(p); }
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
    float rz;                            // one approximate reciprocal (1 ulp), two products: ~2 ulp from the quotients,
    asm("rcp.approx.f32 %0, %1;" : "=f"(rz) : "f"(a.z));   // far inside the 1e-3 contract; no slow path
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

template <bool CG, bool LOOP>
__device__ __forceinline__ void role_normalise(const float4* __restrict__ acc, float* __restrict__ proj,
                                               float* __restrict__ wsum, int32_t* __restrict__ count,
                                               uint8_t* __restrict__ hole, uint32_t* __restrict__ rowmask,
                                               uint32_t* __restrict__ colmask, int* __restrict__ has_holes, int h, int w,
                                               int warp0, int nw, int lane) {
  const int tiles_x = ceil_div(w, 32);
  const int n_strips = 2 * ceil_div(h, 32);          // both halves of every column word get written
  const int n_tasks = tiles_x * n_strips;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // LOOP = false: the stand-alone launch gives every warp exactly one strip (the loop's live range costs the 64-register
  // kernel spills in its inner loop)
  for (int task = warp0; task < n_tasks; task += LOOP ? nw : n_tasks) {
    const int tx = task % tiles_x, strip = task / tiles_x;
    const int x = tx * 32 + lane, y0 = strip * kStrip;
    const bool in_x = x < w;
    const float mx = (x == w - 1) ? 2.0f : 1.0f;
    float4 up = (in_x && y0 > 0 && y0 <= h) ? ldg_f4<CG>(acc + (y0 - 1) * w + x) : zero4;            // cell (y0-1, x)
    float4 up_left = shfl_up1_f4(up);                                                              // cell (y0-1, x-1)
    if (lane == 0) up_left = (x > 0 && y0 > 0 && y0 <= h) ? ldg_f4<CG>(acc + (y0 - 1) * w + x - 1) : zero4;
    uint32_t colbits = 0;
    bool any_hole = false;
#pragma unroll
    for (int r0 = 0; r0 < kStrip; r0 += kBatch) {
      float4 c[kBatch], cl[kBatch];
#pragma unroll
      for (int r = 0; r < kBatch; ++r) {
        const int y = y0 + r0 + r;
        c[r] = (in_x && y < h) ? ldg_f4<CG>(acc + y * w + x) : zero4;
        cl[r] = (lane == 0 && x > 0 && y < h) ? ldg_f4<CG>(acc + y * w + x - 1) : zero4;
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
template <bool CG>
__device__ __forceinline__ void role_fill(const uint32_t* rowmask_, const uint32_t* colmask_,
                                          float* proj, int h, int w, int warp0, int nw, int lane) {
  struct Words { const uint32_t* p; __device__ __forceinline__ uint32_t operator[](int64_t i) const { return CG ? __ldcg(p + i) : __ldg(p + i); } };
  const Words rowmask{rowmask_}, colmask{colmask_};
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
      const float2 q = CG ? __ldcg(pin + yy * w + xx) : pin[yy * w + xx];   // the neighbour's normalised value, as written by the normalise pass
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
          role_normalise<true, true>(a.acc + (b % 3) * P, a.proj + b * P * 2, a.wsum ? a.wsum + b * P : nullptr, a.count + b * P,
                         a.hole + b * P, a.rowmask + b * rw, a.colmask + b * cw, a.flags + 1 + b, a.h, a.w, warp0, nw,
                         lane);
      } else {
        const int b = p - 2;
        if (b >= 0 && b < a.B && *reinterpret_cast<volatile int*>(a.flags + 1 + b) != 0)
          role_fill<true>(a.rowmask + b * rw, a.colmask + b * cw, a.proj + b * P * 2, a.h, a.w, warp0, nw, lane);
      }
    }
    if (p <= a.B) grid.sync();
  }
}

// The same stages as stand-alone kernels, each with its own register budget / occupancy (the merged kernel above runs
// every role at 24 warps per SM; alone the splat keeps 64 in flight): the general path (run_general).
__global__ void __launch_bounds__(kThreads)
stage_splat_kernel(const float* __restrict__ flow, const float* __restrict__ inv_depth, float4* __restrict__ acc, int h, int w) {
  role_splat(flow, inv_depth, acc, h, w, (blockIdx.x * kThreads + threadIdx.x) >> 5, (gridDim.x * kThreads) >> 5, threadIdx.x & 31);
}
__global__ void __launch_bounds__(kThreads, 4)
stage_normalise_kernel(const float4* __restrict__ acc, float* __restrict__ proj, float* __restrict__ wsum,
                       int32_t* __restrict__ count, uint8_t* __restrict__ hole, uint32_t* __restrict__ rowmask,
                       uint32_t* __restrict__ colmask, int* __restrict__ has_holes, int h, int w) {
  role_normalise<false, false>(acc, proj, wsum, count, hole, rowmask, colmask, has_holes, h, w, (blockIdx.x * kThreads + threadIdx.x) >> 5,
                 (gridDim.x * kThreads) >> 5, threadIdx.x & 31);
}
__global__ void __launch_bounds__(kThreads)
stage_fill_kernel(const uint32_t* rowmask, const uint32_t* colmask, const int* has_holes, float* proj, int h, int w) {
  if (*reinterpret_cast<const volatile int*>(has_holes) == 0) return;
  role_fill<false>(rowmask, colmask, proj, h, w, (blockIdx.x * kThreads + threadIdx.x) >> 5, (gridDim.x * kThreads) >> 5,
                   threadIdx.x & 31);
}

// ---------------------------------------------------------------------------------------------
// BOUNDED path: owner-computes tiles in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kTileWarps = 13;                    // 12 rectangle warps + 1 for the tile's left halo column
constexpr int kTileThreads = 32 * kTileWarps;
constexpr int kTileW = 192, kTileH = 64;          // targets per tile = 6 x 2 output sub-blocks of 32 x 32
constexpr int kCellW = kTileW + 1, kCellH = kTileH + 1;   // + the column to the left and the row above
constexpr int kCells = kCellW * kCellH;           // 12 545 cells: 200 720 B of float4 + 12 545 B of claim bytes
constexpr int kMaxBound = 16;
constexpr int kSB = 4;                            // batches of 32 sources handled together (one claim round, 4 RMWs in flight)
constexpr size_t kTileSmem = (size_t)(kCells + 1) * 16 + ((kCells + 1 + 15) / 16) * 16;   // + the dummy cell / claim byte

// A warp's source window is walked row by row, 32 lanes x `n_chunks` per row (64 wide for a bound of 8 = two full
// chunks; the halo-column warp's window is 2*bound+1 wide: one chunk, half the lanes idle, but the same number of
// batches as everybody else).  The batch -> (row, chunk) bookkeeping is warp-uniform and incremental; a lane adds
// its own fixed column.  (ncu on the first versions: per-lane divisions / wrap loops, int->float conversions,
// 64-bit index arithmetic and the branches around predicated shared-memory accesses were most of the instructions.)
struct Walk {
  int n_chunks, n_rows, n_cols;
  int off0;                  // y*w + x of (first row, this lane's column in chunk 0)
  float x0, y0;              // the same position as floats
  int w;
};
struct WalkPos {             // warp-uniform cursor
  int row, chunk;
  __device__ __forceinline__ void step(const Walk& wk) {
    if (++chunk == wk.n_chunks) { chunk = 0; ++row; }
  }
};

struct SrcBatch {       // kSB x 32 sources in registers (invalid sources: flow 0, depth 1)
  float2 f[kSB];
  float d[kSB];
};

__device__ __forceinline__ void load_batch(SrcBatch& s, const Walk& wk, WalkPos& wp, int lane,
                                           const float2* __restrict__ flow, const float* __restrict__ depth) {
#pragma unroll
  for (int u = 0; u < kSB; ++u) {
    const bool ok = wp.row < wk.n_rows && lane + 32 * wp.chunk < wk.n_cols;
    const int off = wk.off0 + wp.row * wk.w + 32 * wp.chunk;
    s.f[u] = make_float2(0.f, 0.f);
    s.d[u] = 1.0f;
    if (ok) {
      s.f[u] = __ldg(flow + off);
      if (depth) s.d[u] = __ldg(depth + off);
    }
    wp.step(wk);
  }
}

// Adds the batch's sources that land in [xa,xb] x [ya,yb] (= this warp's cells, inside the image) to the cells.
// The cells are this warp's alone: plain read-modify-write, no atomics.  Sources of one batch that hit the SAME cell
// take turns: everybody writes its id (sub-batch, lane) to the cell's claim byte, the survivor goes.  Winners
// therefore hold distinct cells, so their four RMWs are independent and overlap.  Everything is branch-free: a source
// with nothing to add (outside the rectangle, or it lost the claim) reads, "updates" and writes a dummy cell.
__device__ __forceinline__ void add_batch(const SrcBatch& cur, const Walk& wk, WalkPos& wp, float4* s_cells,
                                          uint8_t* s_claim, float xa, float xb, float ya, float yb, int cell_org,
                                          int lane, float& vmax) {
  int cell[kSB];
  bool pend[kSB];
  float vx[kSB], vy[kSB];
#pragma unroll
  for (int u = 0; u < kSB; ++u) {
    const float fx = cur.f[u].x, fy = cur.f[u].y;
    const bool ok = wp.row < wk.n_rows && lane + 32 * wp.chunk < wk.n_cols;
    const float x2 = __fadd_rn(wk.x0 + (float)(32 * wp.chunk), fx), y2 = __fadd_rn(wk.y0 + (float)wp.row, fy);
    wp.step(wk);
    vmax = fmaxf(vmax, fmaxf(fabsf(fx), fabsf(fy)));
    pend[u] = ok && x2 >= xa && x2 <= xb && y2 >= ya && y2 <= yb;
    cell[u] = pend[u] ? (int)y2 * kCellW + (int)x2 - cell_org : kCells;     // kCells: the dummy cell / claim byte
    vx[u] = __fmul_rn(-fx, cur.d[u]);
    vy[u] = __fmul_rn(-fy, cur.d[u]);
  }
  bool again;
  int rounds = 0;
  do {
#pragma unroll
    for (int u = 0; u < kSB; ++u) s_claim[cell[u]] = (uint8_t)(u * 32 + lane);
    __syncwarp();
    uint8_t who[kSB];
#pragma unroll
    for (int u = 0; u < kSB; ++u) who[u] = s_claim[cell[u]];
    int idx[kSB];
    float4 c[kSB];
    bool left = false;
#pragma unroll
    for (int u = 0; u < kSB; ++u) {
      const bool win = pend[u] && who[u] == (uint8_t)(u * 32 + lane);
      idx[u] = win ? cell[u] : kCells;
      pend[u] = pend[u] && !win;
      if (!pend[u]) cell[u] = kCells;      // done: from now on this source only touches the dummy (a winner that kept
                                           // writing its id to the real claim byte would starve the others for ever)
      left |= pend[u];
    }
#pragma unroll
    for (int u = 0; u < kSB; ++u) c[u] = s_cells[idx[u]];
#pragma unroll
    for (int u = 0; u < kSB; ++u) {
      c[u].x += vx[u]; c[u].y += vy[u]; c[u].z += cur.d[u]; c[u].w += 1.0f;
      s_cells[idx[u]] = c[u];
    }
    again = __any_sync(0xffffffffu, left);
    __syncwarp();
    // every round retires at least one source per contested cell: at most kSB * 32 rounds.  The guard turns a logic
    // error into a redo by the general path instead of a hung GPU.
    if (++rounds > kSB * 32 + 2) { vmax = __int_as_float(0x7f800000); break; }
  } while (again);
}

__global__ void __launch_bounds__(kTileThreads, 1)
projection_tiled_kernel(const ProjArgs a) {
  extern __shared__ float4 s_cells[];
  uint8_t* s_claim = reinterpret_cast<uint8_t*>(s_cells + kCells + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = a.h, w = a.w, D = a.bound;
  const int tiles_x = ceil_div(w, kTileW), tiles_y = ceil_div(h, kTileH);
  const int n_tiles = tiles_x * tiles_y * a.B;
  const int P = h * w;                            // <= 2^30
  const float xmax = (float)(w - 1), ymax = (float)(h - 1);
  const int rw_x = ceil_div(w, 32);
  const int rw = h * rw_x, cw = ceil_div(h, 32) * w;
  // accumulation rectangle of this warp inside the 193 x 65 cell tile: warps 0-11 take 4 x 3 rectangles of 48 x
  // 22(21) cells over local columns 1..192, warp 12 the left halo column (local column 0, all 65 rows)
  const bool halo = warp == 12;
  const int bx = warp & 3, by = warp >> 2;
  const int cx_lo = halo ? 0 : 1 + 48 * bx, cx_hi = halo ? 1 : 49 + 48 * bx;
  const int cy_lo = halo ? 0 : 22 * by, cy_hi = halo ? kCellH : (by == 2 ? kCellH : 22 * (by + 1));
  // output sub-block of warps 0-11: 32 x 32 targets
  const int ox = (warp % 6) * 32, oy = (warp / 6) * 32;
  float vmax = 0.0f;                              // largest |flow component| this thread has looked at

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tb = tile / (tiles_x * tiles_y);
    const int tr = tile - tb * tiles_x * tiles_y;
    const int tx0 = (tr % tiles_x) * kTileW, ty0 = (tr / tiles_x) * kTileH;   // first target of the tile
    const float2* flow = reinterpret_cast<const float2*>(a.flow) + (int64_t)tb * P;
    const float* depth = a.inv_depth ? a.inv_depth + (int64_t)tb * P : nullptr;

    for (int i = threadIdx.x; i < kCells; i += kTileThreads) s_cells[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    // ---- accumulate: sources of the window [rectangle - D, rectangle + D]; global cell = local + (ty0-1, tx0-1)
    {
      const int gcx_lo = tx0 - 1 + cx_lo, gcx_hi = tx0 - 1 + cx_hi;       // this warp's cells, global, hi exclusive
      const int gcy_lo = ty0 - 1 + cy_lo, gcy_hi = ty0 - 1 + cy_hi;
      const int gx_lo = max(gcx_lo - D, 0), gx_hi = min(gcx_hi - 1 + D, w - 1);   // source window, inclusive
      const int gy_lo = max(gcy_lo - D, 0), gy_hi = min(gcy_hi - 1 + D, h - 1);
      const int Ws = gx_hi - gx_lo + 1, Hs = gy_hi - gy_lo + 1;
      const bool any = Ws > 0 && Hs > 0 && gcx_hi > 0 && gcy_hi > 0;
      // "lands in one of my cells and inside the image" as four float compares on the target position (x2, y2):
      // int(x2) in [lo, hi) <=> lo <= x2 < hi for x2 >= 0; `x2 < hi` is `x2 <= pred(hi)`, and at the image edge the
      // bound is Appendix B's x2 <= w-1.  NaN fails every compare.
      const float xa = (float)max(gcx_lo, 0), xb = gcx_hi <= w - 1 ? nextafterf((float)gcx_hi, 0.0f) : xmax;
      const float ya = (float)max(gcy_lo, 0), yb = gcy_hi <= h - 1 ? nextafterf((float)gcy_hi, 0.0f) : ymax;
      const int cell_org = (ty0 - 1) * kCellW + (tx0 - 1);
      Walk wk;
      wk.n_cols = Ws;
      wk.n_rows = any ? Hs : 0;
      wk.n_chunks = max(ceil_div(Ws, 32), 1);
      wk.off0 = gy_lo * w + gx_lo + lane;
      wk.x0 = (float)(gx_lo + lane);
      wk.y0 = (float)gy_lo;
      wk.w = w;
      const int nb = wk.n_rows * wk.n_chunks;
      // three register buffers: the loads of the next two batches fly under the arithmetic of the current one
      // (ncu: with one batch of look-ahead a fifth of the stall samples still sat on the first use of a load)
      SrcBatch b0, b1, b2;
      WalkPos lp = {0, 0}, ap = {0, 0};          // load cursor, add cursor
      if (nb > 0) load_batch(b0, wk, lp, lane, flow, depth);
      if (nb > kSB) load_batch(b1, wk, lp, lane, flow, depth);
      for (int j = 0; j < nb; j += 3 * kSB) {
        if (j + 2 * kSB < nb) load_batch(b2, wk, lp, lane, flow, depth);
        add_batch(b0, wk, ap, s_cells, s_claim, xa, xb, ya, yb, cell_org, lane, vmax);
        if (j + kSB >= nb) break;
        if (j + 3 * kSB < nb) load_batch(b0, wk, lp, lane, flow, depth);
        add_batch(b1, wk, ap, s_cells, s_claim, xa, xb, ya, yb, cell_org, lane, vmax);
        if (j + 2 * kSB >= nb) break;
        if (j + 4 * kSB < nb) load_batch(b1, wk, lp, lane, flow, depth);
        add_batch(b2, wk, ap, s_cells, s_claim, xa, xb, ya, yb, cell_org, lane, vmax);
      }
    }
    __syncthreads();

    // ---- output: 32 x 32 targets per warp, lane = column, rows walked; local cell of target (ty,tx) = (ty+1-ty0, tx+1-tx0)
    if (!halo && ty0 + oy < h && tx0 + ox < w) {
      const int x = tx0 + ox + lane;
      const bool in_x = x < w;
      const float mx = (x == w - 1) ? 2.0f : 1.0f;
      const float4* col = s_cells + oy * kCellW + ox + lane + 1;     // cell (row above the first target, this column)
      float4 up = col[0];
      float4 up_left = col[-1];
      uint32_t colbits = 0;
      bool any_hole = false;
      const int64_t img = (int64_t)tb * P;
      float2* proj = reinterpret_cast<float2*>(a.proj) + img;
      float* wsum = a.wsum ? a.wsum + img : nullptr;
      int32_t* count = a.count + img;
      uint8_t* hole = a.hole + img;
      uint32_t* rowm = a.rowmask + (int64_t)tb * rw + ((tx0 + ox) >> 5);
      const int rows = min(32, h - (ty0 + oy));
      int p = (ty0 + oy) * w + x;
      for (int r = 0; r < rows; ++r, p += w) {
        col += kCellW;
        const float4 c11 = col[0];
        float4 c10 = shfl_up1_f4(c11);
        if (lane == 0) c10 = col[-1];
        const float4 c01 = up, c00 = up_left;
        up = c11;
        up_left = c10;
        const float my = (ty0 + oy + r == h - 1) ? 2.0f : 1.0f;
        float4 t;
        t.x = fmaf(fmaf(c11.x, mx, c10.x), my, fmaf(c01.x, mx, c00.x));
        t.y = fmaf(fmaf(c11.y, mx, c10.y), my, fmaf(c01.y, mx, c00.y));
        t.z = fmaf(fmaf(c11.z, mx, c10.z), my, fmaf(c01.z, mx, c00.z));
        t.w = fmaf(fmaf(c11.w, mx, c10.w), my, fmaf(c01.w, mx, c00.w));
        bool is_hole = false;
        if (in_x) {
          float2 o;
          is_hole = normalise_target(t, o);
          proj[p] = o;
          if (wsum) wsum[p] = is_hole ? 0.0f : t.z;
          count[p] = (int32_t)t.w;
          hole[p] = is_hole ? 1 : 0;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, in_x && !is_hole);
        if (lane == 0) rowm[(ty0 + oy + r) * rw_x] = m;
        colbits |= (uint32_t)(in_x && !is_hole) << r;
        any_hole |= in_x && is_hole;
      }
      if (in_x) a.colmask[(int64_t)tb * cw + ((ty0 + oy) >> 5) * w + x] = colbits;
      if (__any_sync(0xffffffffu, any_hole) && lane == 0) *reinterpret_cast<volatile int*>(a.flags + 1 + tb) = 1;
    }
    __syncthreads();
  }
  // a source beyond the promised bound: the windows above may have missed cells it reaches -> redo by the general path
  if (__any_sync(0xffffffffu, vmax > (float)D) && lane == 0) *reinterpret_cast<volatile int*>(a.flags) = 1;
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
      role_fill<false>(a.rowmask + b * rw, a.colmask + b * cw, a.proj + b * P * 2, a.h, a.w, warp0, nw, lane);
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

// General path: per image memset -> splat -> normalise -> fill on the caller's stream, one L2-resident cell array.
// Measured alternatives on B200 (1080p, batch 8, us per image, smooth +-8 / i.i.d. +-64 / occlusion; profiles/
// time_ops_r02*.log): this sequence 45 / 59 / 45; images on 2 internal streams with 2 cell arrays 50 / 88 / 49; on 3
// streams 52 / 133 / 51 -- more than one 33 MB cell array in flight falls out of L2 under the streaming traffic and the
// scattered reductions go to DRAM; the whole batch pipelined inside one cooperative kernel (projection_pipeline_kernel
// above, kept as the fallback of the bounded path) 69 / 255 / 103 -- every role then runs at the merged kernel's 24
// warps per SM.
int run_general(const ProjArgs& a, cudaStream_t st) {
  const int h = a.h, w = a.w;
  const int64_t P = (int64_t)h * w;
  const int64_t rw = (int64_t)h * ceil_div(w, 32), cw = (int64_t)ceil_div(h, 32) * w;
  cudaError_t e = cudaMemsetAsync(a.flags + 1, 0, (size_t)a.B * 4, st);
  if (e != cudaSuccess) return cuda_status(e);
  const int n_tasks = ceil_div(w, 32) * ceil_div(h, kRows);
  const int splat_blocks = ceil_div(n_tasks, kThreads / 32);
  const int norm_blocks = ceil_div(ceil_div(w, 32) * 2 * ceil_div(h, 32), kThreads / 32);
  const int fill_blocks = ceil_div(h * ceil_div(w, 32), kThreads / 32);   // one warp per 32-pixel row word
  for (int b = 0; b < a.B; ++b) {
    if ((e = cudaMemsetAsync(a.acc, 0, (size_t)P * sizeof(float4), st)) != cudaSuccess) return cuda_status(e);
    stage_splat_kernel<<<splat_blocks, kThreads, 0, st>>>(a.flow + b * P * 2, a.inv_depth ? a.inv_depth + b * P : nullptr, a.acc, h, w);
    int rc = after_launch();
    if (rc) return rc;
    stage_normalise_kernel<<<norm_blocks, kThreads, 0, st>>>(a.acc, a.proj + b * P * 2, a.wsum ? a.wsum + b * P : nullptr,
                                                            a.count + b * P, a.hole + b * P, a.rowmask + b * rw,
                                                            a.colmask + b * cw, a.flags + 1 + b, h, w);
    if ((rc = after_launch())) return rc;
    stage_fill_kernel<<<fill_blocks, kThreads, 0, st>>>(a.rowmask + b * rw, a.colmask + b * cw, a.flags + 1 + b,
                                                       a.proj + b * P * 2, h, w);
    if ((rc = after_launch())) return rc;
  }
  return VSR_OK;
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
  if (!bounded) return run_general(a, st);

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
