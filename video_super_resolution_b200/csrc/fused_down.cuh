// fused_down.cuh -- one kernel for the HR half of a feedback group:
//
//     lr[i+1] (pre-activation sums) = Conv8x8s4( PReLU( Conv1x1( cat(hr[0..i]) ) ) )
//     ref: FeedbackBlock.forward, SRProjectionModule.py:70-80 (downtranBlocks[i-1] then downBlocks[i]),
//          intended dense-concat dataflow
//
// Layer by layer this is the most expensive part of the stack in HBM bytes: the 1x1 "downtran" reads
// (i+1) HR feature maps and writes one, the 8x8-s4 conv reads it back.  Here the intermediate HR map
// never leaves the SM:
//
//   tile   = 16 x 8 blocks of the HR block layout (128 blocks = 128 UMMA rows) of one map
//   phase A (per group g of 4 sub-positions): D_A[128, 4x32] = sum_j hr_j[tile, s, :] * Wt_j^T
//            A operands: TMA boxes [32 ch, 1 s, 16, 8] of the 5-D view (c, s, Xb, Yb, map); N = 32
//   convert: epilogue warps read D_A from TMEM, add bias, PReLU, zero the padding ring, round to
//            BF16 and write the [128, 128] tile into shared memory in the 128B-swizzled K-major
//            layout the tensor core reads (fence.proxy.async), i.e. they PRODUCE phase B's A operand
//   phase B: D_B[128, 4 taps x 32] += H_g[128, 128] * Wd_g[128, 128]^T        (K = 512 over 4 groups)
//            "output-shift" form of the strided conv: block (Yb,Xb) contributes its tap-(dy,dx)
//            partial product to LR pixel (Yb-dy, Xb-dx); every block is read exactly once
//   epilogue: the four 32-channel partials of each block are added into an FP32 LR accumulator
//            with vector reductions (red.global.add.v4.f32); finalize_lr_kernel applies
//            bias + PReLU, rounds to BF16 and re-zeroes the accumulator.
//
// HAS_TRAN = false is group 0 (no downtran: H = hr[0]); phase B's A operand then comes straight from
// TMA.  Warp roles as in igemm.cuh: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#pragma once
#include "igemm.cuh"

namespace vsr {

constexpr int kFusedThreads = 192;
constexpr int kFusedMaxStages = 8;
constexpr int kWdBytes = 128 * 512 * 2;      // resident 8x8-s4 weights, 8 chunks [128 n][64 k]
constexpr int kWtChunkBytes = 32 * 32 * 2;   // one 32x32 downtran slice
constexpr int kHBytes = 128 * 128 * 2;       // phase-B A operand of one sub-position group

struct alignas(64) FusedDownParams {
  CUtensorMap hr_maps[kMaxSources];   // HAS_TRAN: 5-D (32, 16, w+1, h+1, B), box (32,1,16,8,1), 64B swizzle
  CUtensorMap h0_map;                 // !HAS_TRAN: 4-D (512, w+1, h+1, B), box (64,16,8,1), 128B swizzle
  CUtensorMap wt_map;                 // 2-D (32*nsrc, 32), box (32,32), 64B swizzle
  CUtensorMap wd_map;                 // 2-D (512, 128), box (64,128), 128B swizzle
  int32_t nsrc;
  int32_t num_stages;
  int32_t tiles_x, tiles_y, batch;
  int32_t lr_h, lr_w;
  const float* tran_bias;             // [32] biases + [1] PReLU slope of the downtran
  float* acc;                         // (B, h, w, 32) fp32, zero on entry
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool HAS_TRAN>
inline size_t fused_down_smem_bytes(int nsrc, int num_stages) {
  size_t wt = HAS_TRAN ? ((size_t)nsrc * kWtChunkBytes + 1023) / 1024 * 1024 : 0;
  size_t stage = HAS_TRAN ? 8192 : 2 * 16384;
  return 1024 + kWdBytes + wt + (HAS_TRAN ? kHBytes : 0) + (size_t)num_stages * stage + 1024;
}

template <bool HAS_TRAN>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_down_kernel(const __grid_constant__ FusedDownParams p) {
  constexpr int kStageBytes = HAS_TRAN ? 8192 : 2 * 16384;
  constexpr int kTmemCols = HAS_TRAN ? 512 : 256;
  constexpr uint32_t kDB = HAS_TRAN ? 256 : 0;       // first column of the two D_B accumulators

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_wd = smem;
  uint8_t* s_wt = s_wd + kWdBytes;
  const int wt_region = HAS_TRAN ? ((p.nsrc * kWtChunkBytes + 1023) / 1024 * 1024) : 0;
  uint8_t* s_h = s_wt + wt_region;
  uint8_t* s_a = s_h + (HAS_TRAN ? kHBytes : 0);
  uint8_t* tail = s_a + p.num_stages * kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);      // [kFusedMaxStages]
  uint64_t* empty_bar = full_bar + kFusedMaxStages;            // [kFusedMaxStages]
  uint64_t* w_full = empty_bar + kFusedMaxStages;              // [1]
  uint64_t* da_full = w_full + 1;                              // [2]
  uint64_t* da_empty = da_full + 2;                            // [2]
  uint64_t* h_full = da_empty + 2;                             // [1]
  uint64_t* h_empty = h_full + 1;                              // [1]
  uint64_t* db_full = h_empty + 1;                             // [2]
  uint64_t* db_empty = db_full + 2;                            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(db_empty + 2);
  float* s_bias = reinterpret_cast<float*>(tail + 512);        // 33 floats

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&da_full[s], 1);
      mbar_init(&da_empty[s], 128);
      mbar_init(&db_full[s], 1);
      mbar_init(&db_empty[s], 128);
    }
    mbar_init(h_full, 128);
    mbar_init(h_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr);
  if (HAS_TRAN)
    for (int i = threadIdx.x; i < 33; i += kFusedThreads) s_bias[i] = p.tran_bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto tile_coord = [&](int tile, int& x0, int& y0, int& b) {
    x0 = (tile % p.tiles_x) * 16;
    int r = tile / p.tiles_x;
    y0 = (r % p.tiles_y) * 8;
    b = r / p.tiles_y;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)(kWdBytes + (HAS_TRAN ? p.nsrc * kWtChunkBytes : 0)));
      for (int kc = 0; kc < 8; ++kc) tma_load_2d(s_wd + kc * 16384, &p.wd_map, w_full, kc * 64, 0);
      if (HAS_TRAN)
        for (int j = 0; j < p.nsrc; ++j) tma_load_2d(s_wt + j * kWtChunkBytes, &p.wt_map, w_full, j * 32, 0);
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int x0, y0, b;
        tile_coord(tile, x0, y0, b);
        if (HAS_TRAN) {
          for (int sp = 0; sp < 16; ++sp)
            for (int j = 0; j < p.nsrc; ++j) {
              mbar_wait(&empty_bar[s], phase ^ 1);
              mbar_expect_tx(&full_bar[s], kStageBytes);
              tma_load_5d(s_a + s * kStageBytes, &p.hr_maps[j], &full_bar[s], 0, sp, x0, y0, b);
              if (++s == p.num_stages) { s = 0; phase ^= 1; }
            }
        } else {
          for (int g = 0; g < 4; ++g) {
            mbar_wait(&empty_bar[s], phase ^ 1);
            mbar_expect_tx(&full_bar[s], kStageBytes);
            tma_load_4d(s_a + s * kStageBytes, &p.h0_map, &full_bar[s], g * 128, x0, y0, b);
            tma_load_4d(s_a + s * kStageBytes + 16384, &p.h0_map, &full_bar[s], g * 128 + 64, x0, y0, b);
            if (++s == p.num_stages) { s = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_a = make_idesc(32);
    constexpr uint32_t idesc_b = make_idesc(128);
    mbar_wait(w_full, 0);
    tc_fence_after();
    int s = 0;
    uint32_t phase = 0;
    uint32_t n_a[2] = {0, 0}, n_b[2] = {0, 0}, n_h = 0;
    int tb = 0;

    // phase A of sub-position group g: D_A[g&1] = sum_j A(s, j) * Wt_j^T for the 4 sub-positions
    auto issue_a = [&](int g) {
      const int buf = g & 1;
      mbar_wait(&da_empty[buf], (n_a[buf] & 1) ^ 1);
      tc_fence_after();
      for (int sl = 0; sl < 4; ++sl)
        for (int j = 0; j < p.nsrc; ++j) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(s_a + s * kStageBytes);
            const uint32_t b_addr = smem_u32(s_wt + j * kWtChunkBytes);
            const uint32_t d = tmem_base + (uint32_t)(buf * 128 + sl * 32);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d, make_smem_desc<64>(a_addr + k * 32), make_smem_desc<64>(b_addr + k * 32), idesc_a,
                        (uint32_t)((j | k) != 0));
            umma_commit(&empty_bar[s]);
          }
          __syncwarp();
          if (++s == p.num_stages) { s = 0; phase ^= 1; }
        }
      if (lane == 0) umma_commit(&da_full[buf]);
      __syncwarp();
      ++n_a[buf];
    };
    // phase B of group g: D_B[tb] += H_g * Wd_g^T
    auto issue_b = [&](int g) {
      uint32_t a_addr;
      if (HAS_TRAN) {
        mbar_wait(h_full, n_h & 1);
        a_addr = smem_u32(s_h);
      } else {
        mbar_wait(&full_bar[s], phase);
        a_addr = smem_u32(s_a + s * kStageBytes);
      }
      if (g == 0) mbar_wait(&db_empty[tb], (n_b[tb] & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t d = tmem_base + kDB + (uint32_t)(tb * 128);
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          const uint32_t b_addr = smem_u32(s_wd + (g * 2 + kc) * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d, make_smem_desc<128>(a_addr + kc * 16384 + k * 32), make_smem_desc<128>(b_addr + k * 32),
                      idesc_b, (uint32_t)((g | kc | k) != 0));
        }
        if (HAS_TRAN) umma_commit(h_empty);
        else umma_commit(&empty_bar[s]);
        if (g == 3) umma_commit(&db_full[tb]);
      }
      __syncwarp();
      if (HAS_TRAN) ++n_h;
      else if (++s == p.num_stages) { s = 0; phase ^= 1; }
      if (g == 3) { ++n_b[tb]; tb ^= 1; }
    };

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (HAS_TRAN) {
        issue_a(0);
        issue_a(1);
        issue_b(0);
        issue_a(2);
        issue_b(1);
        issue_a(3);
        issue_b(2);
        issue_b(3);
      } else {
        for (int g = 0; g < 4; ++g) issue_b(g);
      }
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t m_a[2] = {0, 0}, m_b[2] = {0, 0}, m_h = 0;
    int tb = 0;
    const float slope = HAS_TRAN ? s_bias[32] : 0.0f;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int x0, y0, b;
      tile_coord(tile, x0, y0, b);
      const int Yb = y0 + (row >> 4), Xb = x0 + (row & 15);
      const bool in_tensor = (Yb <= p.lr_h) && (Xb <= p.lr_w);
      if (HAS_TRAN) {
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          const int buf = g & 1;
          mbar_wait(&da_full[buf], m_a[buf] & 1);
          tc_fence_after();
          uint32_t o[4][16];
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            uint32_t v[32];
            tmem_ld32(lane_base + (uint32_t)(buf * 128 + sl * 32), v);
            tmem_ld_wait();
            const int s16 = g * 4 + sl, ry = s16 >> 2, rx = s16 & 3;
            const bool ring = (Yb == 0 && ry < 2) || (Yb == p.lr_h && ry >= 2) || (Xb == 0 && rx < 2) ||
                              (Xb == p.lr_w && rx >= 2);
            const bool keep = in_tensor && !ring;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a = prelu(__uint_as_float(v[2 * j]) + s_bias[2 * j], slope, 1);
              float c = prelu(__uint_as_float(v[2 * j + 1]) + s_bias[2 * j + 1], slope, 1);
              o[sl][j] = keep ? pack_bf16(a, c) : 0u;
            }
          }
          tc_fence_before();
          mbar_arrive(&da_empty[buf]);
          ++m_a[buf];
          mbar_wait(h_empty, (m_h & 1) ^ 1);
          // K index inside the group: k = sl*32 + c -> 64-element chunk kc = sl>>1, 16-byte piece (sl&1)*4 + j
          uint8_t* hrow = s_h + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
          for (int sl = 0; sl < 4; ++sl)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int piece = (sl & 1) * 4 + j;
              *reinterpret_cast<uint4*>(hrow + (sl >> 1) * 16384 + ((piece ^ (row & 7)) << 4)) =
                  make_uint4(o[sl][4 * j], o[sl][4 * j + 1], o[sl][4 * j + 2], o[sl][4 * j + 3]);
            }
          fence_proxy_async_smem();
          mbar_arrive(h_full);
          ++m_h;
        }
      }
      // final: D_B[tb] -> shifted accumulation into the LR sums
      mbar_wait(&db_full[tb], m_b[tb] & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 4; ++t) {
        uint32_t v[32];
        tmem_ld32(lane_base + kDB + (uint32_t)(tb * 128 + t * 32), v);
        tmem_ld_wait();
        if (t == 3) {
          tc_fence_before();
          mbar_arrive(&db_empty[tb]);
        }
        const int Y = Yb - (t >> 1), X = Xb - (t & 1);
        if (in_tensor && Y >= 0 && Y < p.lr_h && X >= 0 && X < p.lr_w) {
          float* dst = p.acc + (((int64_t)b * p.lr_h + Y) * p.lr_w + X) * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            red_add_v4(dst + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
      ++m_b[tb];
      tb ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// lr_out[p, c] = bf16(PReLU(acc[p, c] + bias[c])); acc[p, c] = 0.   8 channels per thread.
__global__ void __launch_bounds__(256)
finalize_lr_kernel(float4* __restrict__ acc, const float* __restrict__ bias, uint4* __restrict__ out, int64_t n8) {
  const float slope = __ldg(bias + 32);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 3) * 8;
    float4 a = acc[2 * i], b = acc[2 * i + 1];
    acc[2 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = prelu(r[k] + __ldg(bias + c + k), slope, 1);
    out[i] = make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]), pack_bf16(r[4], r[5]), pack_bf16(r[6], r[7]));
  }
}

}  // namespace vsr
