// bf16 GEMM, CTA-pair variant: two SMs of a cluster cooperate on one 256×256 output tile with
// tcgen05.mma.cta_group::2 (UMMA M=256, N=256, K=16).  Each CTA stages its own 128 rows of A and HALF of the B tile
// (128 of the 256 N-rows), so per-CTA shared-memory fill and L2→SMEM traffic for B halve versus the single-CTA kernel
// and the ring deepens to 6 stages × 32 KB.  The leader CTA (cluster rank 0) issues every MMA; completion is
// multicast to both CTAs' barriers with tcgen05.commit…multicast::cluster; each CTA drains its own 128 TMEM lanes.
//
//   warp 0  TMA producer (both CTAs; cta_group::2 loads signal the leader's full barrier)
//   warp 1  MMA issuer   (leader only) + TMEM allocator (both, cta_group::2)
//   warps 2-5 epilogue   (both)
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace nxd {
namespace g2 {

constexpr int TILE_M = 256;         // per pair
constexpr int CTA_M = 128;
constexpr int TILE_N = 256;
constexpr int HALF_N = 128;         // B rows staged per CTA
constexpr int BK = 64;
constexpr int UK = 16;
constexpr int kStages = 6;
constexpr int kAcc = 2;
constexpr int kABytes = CTA_M * BK * 2;      // 16 KB
constexpr int kBBytes = HALF_N * BK * 2;     // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kThreads = 192;
constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
constexpr int kTmemCols = kAcc * TILE_N;     // 512

template <typename OutT>
NXD_DEVICE void store_chunk(OutT* orow, int col0, int N, const uint32_t (&r)[32], int accumulate) {
  if constexpr (sizeof(OutT) == 2) {
    if (col0 + 32 <= N) {
      uint4 pk[4];
      __nv_bfloat162* h = (__nv_bfloat162*)pk;
      if (accumulate) {
        const uint4* old = (const uint4*)(orow + col0);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint4 o = old[v];
          const __nv_bfloat162* oh = (const __nv_bfloat162*)&o;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 of = __bfloat1622float2(oh[j]);
            h[v * 4 + j] = __floats2bfloat162_rn(__uint_as_float(r[v * 8 + 2 * j]) + of.x,
                                                 __uint_as_float(r[v * 8 + 2 * j + 1]) + of.y);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) h[j] = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      }
      uint4* dst = (uint4*)(orow + col0);
#pragma unroll
      for (int v = 0; v < 4; ++v) dst[v] = pk[v];
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < N) {
          float v = __uint_as_float(r[j]);
          if (accumulate) v += __bfloat162float(((__nv_bfloat16*)orow)[col0 + j]);
          ((__nv_bfloat16*)orow)[col0 + j] = __float2bfloat16_rn(v);
        }
      }
    }
  } else {
    if (col0 + 32 <= N) {
      float4* dst = (float4*)(orow + col0);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        float4 o = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                               __uint_as_float(r[4 * v + 3]));
        if (accumulate) { const float4 p = dst[v]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
        dst[v] = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < N) {
          float v = __uint_as_float(r[j]);
          if (accumulate) v += ((float*)orow)[col0 + j];
          ((float*)orow)[col0 + j] = v;
        }
      }
    }
  }
}

NXD_DEVICE void tile_coords(int tile, int tiles_m, int tiles_n, int& m_blk, int& n_blk) {
  constexpr int GROUP = 4;  // 4 × 256 rows per raster group
  const int per_group = GROUP * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * GROUP;
  const int gsz = min(GROUP, tiles_m - first_m);
  const int in = tile - g * per_group;
  m_blk = first_m + in % gsz;
  n_blk = in / gsz;
}

template <bool A_KMAJOR, bool B_KMAJOR, typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      OutT* __restrict__ out, int M, int N, int K, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 2 * kAcc);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + kAcc);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int tiles_m = (M + TILE_M - 1) / TILE_M, tiles_n = (N + TILE_N - 1) / TILE_N;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 2);     // leader: own arrive.expect_tx + peer's remote arrive
      mbar_init(bar_empty + 8 * i, 1);    // one multicast tcgen05.commit
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 2 * 128);   // epilogue threads of both CTAs (used on the leader)
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) { tcgen05_alloc_2cta(smem_u32(tmem_slot), kTmemCols); tcgen05_relinquish_2cta(); }
  tcgen05_fence_before();
  cluster_sync_all();               // barriers of both CTAs are initialised before any remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, m_blk, n_blk);
        const int m0 = m_blk * TILE_M + (int)cta * CTA_M;
        const int n0 = n_blk * TILE_N + (int)cta * HALF_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;          // same offset in the leader's smem
          if (leader) mbar_expect_tx(full, 2 * kStageBytes);
          else mbar_arrive_cluster(mapa_shared(full, 0));
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
          const int k0 = kb * BK;
          if constexpr (A_KMAJOR) {
            tma_load_2d_2cta(sa, &tma_a, full, k0, m0);                          // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < CTA_M / 64; ++j) tma_load_2d_2cta(sa + j * 8192, &tma_a, full, m0 + j * 64, k0);
          }
          if constexpr (B_KMAJOR) {
            tma_load_2d_2cta(sb, &tma_b, full, k0, n0);                          // box {64 k, 128 n}
          } else {
#pragma unroll
            for (int j = 0; j < HALF_N / 64; ++j) tma_load_2d_2cta(sb + j * 8192, &tma_b, full, n0 + j * 64, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(!A_KMAJOR, !B_KMAJOR, TILE_M, TILE_N);
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * TILE_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = A_KMAJOR ? make_smem_desc(sa + k * 32, 16, 1024) : make_smem_desc(sa + k * 2048, 8192, 1024);
            const uint64_t db = B_KMAJOR ? make_smem_desc(sb + k * 32, 16, 1024) : make_smem_desc(sb + k * 2048, 8192, 1024);
            tcgen05_mma_f16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tcgen05_commit_2cta(bar_empty + 8 * stage, 0b11);    // frees the stage in BOTH CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_2cta(bar_tfull + 8 * as, 0b11);         // accumulators ready in both CTAs
        if (++as == kAcc) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (both CTAs; own 128 TMEM lanes) =====
    const int q = warp & 3;
    const uint32_t tempty_leader = mapa_shared(bar_tempty, 0);
    int as = 0; uint32_t aphase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, m_blk, n_blk);
      mbar_wait(bar_tfull + 8 * as, aphase);
      tcgen05_fence_after();
      const int row = m_blk * TILE_M + (int)cta * CTA_M + q * 32 + lane;
      const int n0 = n_blk * TILE_N;
      OutT* orow = out + (size_t)row * N;
      const bool row_ok = row < M;
#pragma unroll 1
      for (int c = 0; c < TILE_N / 32; ++c) {
        uint32_t r[32];
        tcgen05_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * TILE_N + c * 32, r);
        tcgen05_wait_ld();
        const int col0 = n0 + c * 32;
        if (row_ok && col0 < N) store_chunk<OutT>(orow, col0, N, r, accumulate);
      }
      tcgen05_fence_before();
      mbar_arrive_cluster(tempty_leader + 8 * as);   // leader's barrier collects both CTAs' epilogues
      if (++as == kAcc) { as = 0; aphase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();               // peer must be done with our smem/TMEM before teardown
  if (warp == 1) { __syncwarp(); tcgen05_dealloc_2cta(tmem_base, kTmemCols); }
}

}  // namespace g2

// ------------------------------------------------------------------ host side
CUtensorMap make_tmap_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);
int device_sm_count();

template <bool AK, bool BK_, typename OutT>
static void launch2(const CUtensorMap& ta, const CUtensorMap& tb, void* out, int M, int N, int K, bool accumulate, int grid,
                    cudaStream_t st) {
  auto kern = g2::gemm_bf16_2cta_kernel<AK, BK_, OutT>;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::kSmem));
    configured = true;
  }
  kern<<<grid, g2::kThreads, g2::kSmem, st>>>(ta, tb, (OutT*)out, M, N, K, accumulate ? 1 : 0);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void gemm_bf16_2cta(const void* a, const void* b, void* out, int M, int N, int K, bool trans_a, bool trans_b, int out_dt,
                    bool accumulate, cudaStream_t st) {
  const bool AK = !trans_a, BK_ = trans_b;
  const CUtensorMap ta = AK ? make_tmap_bf16(a, M, K, g2::BK, g2::CTA_M) : make_tmap_bf16(a, K, M, 64, g2::BK);
  const CUtensorMap tb = BK_ ? make_tmap_bf16(b, N, K, g2::BK, g2::HALF_N) : make_tmap_bf16(b, K, N, 64, g2::BK);
  const int tiles = ((M + g2::TILE_M - 1) / g2::TILE_M) * ((N + g2::TILE_N - 1) / g2::TILE_N);
  const int pairs = device_sm_count() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
#define NXD_L2(AKv, BKv)                                                                          \
  do {                                                                                            \
    if (out_dt == kBF16) launch2<AKv, BKv, __nv_bfloat16>(ta, tb, out, M, N, K, accumulate, grid, st); \
    else launch2<AKv, BKv, float>(ta, tb, out, M, N, K, accumulate, grid, st);                   \
  } while (0)
  if (AK && BK_) NXD_L2(true, true);
  else if (AK && !BK_) NXD_L2(true, false);
  else if (!AK && !BK_) NXD_L2(false, false);
  else NXD_L2(false, true);
#undef NXD_L2
}

}  // namespace nxd
