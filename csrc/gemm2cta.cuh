// Shared pieces of the CTA-pair (cta_group::2) bf16 GEMM kernels: tile constants, the TMEM→global epilogue store and the
// 256-row raster.  Included by gemm2cta_sm100.cu (plain + unicast TP kernels) and tp_nvls_sm100.cu (NVLS TP kernels).
#pragma once
#include <cuda.h>

#include "common.cuh"
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
// epilogue staging: two 128-row x 64-column bf16 slabs (128B-swizzled rows), written by the epilogue warps and drained by
// TMA bulk tensor stores — coalesced, asynchronous, no per-thread row-strided global stores
constexpr int kEpiSlabCols = 64;
constexpr int kEpiSlabBytes = CTA_M * kEpiSlabCols * 2;   // 16 KB
constexpr int kEpiBufs = 2;
constexpr int kEpiOffset = kStages * kStageBytes;
constexpr int kBarOffset = kEpiOffset + kEpiBufs * kEpiSlabBytes;
constexpr int kSmem = kBarOffset + 1024 + 256;
constexpr int kTmemCols = kAcc * TILE_N;     // 512
static_assert(kSmem <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");

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
      if (accumulate == 2) {                 // split-K partial: vector atomic add (fire-and-forget reduction at L2)
#pragma unroll
        for (int v = 0; v < 8; ++v)
          asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                       ::"l"(dst + v), "f"(__uint_as_float(r[4 * v])), "f"(__uint_as_float(r[4 * v + 1])),
                         "f"(__uint_as_float(r[4 * v + 2])), "f"(__uint_as_float(r[4 * v + 3])) : "memory");
      } else {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          float4 o = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                 __uint_as_float(r[4 * v + 3]));
          if (accumulate) { const float4 p = dst[v]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
          dst[v] = o;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < N) {
          float v = __uint_as_float(r[j]);
          if (accumulate == 2) { atomicAdd(((float*)orow) + col0 + j, v); continue; }
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


NXD_DEVICE void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
NXD_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_> NXD_DEVICE void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_) : "memory"); }
template <int N_> NXD_DEVICE void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N_) : "memory"); }
NXD_DEVICE void bulk_wait_n(int n) {       // all but the newest n bulk groups have fully completed (n in 0..4)
  switch (n) {
    case 0: bulk_wait<0>(); break;
    case 1: bulk_wait<1>(); break;
    case 2: bulk_wait<2>(); break;
    case 3: bulk_wait<3>(); break;
    default: bulk_wait<4>(); break;
  }
}

// Epilogue of one 128x256 accumulator (this CTA's TMEM lanes) to a row-major bf16 matrix through TMA stores.
// Executed by the 128 epilogue threads (warps 2-5, named barrier 1); thread `issuer` (one of them) owns the bulk groups.
// Per 64-column slab: wait until the staging buffer's previous store has read it, tcgen05.ld 2x32 columns, pack to bf16,
// write the thread's 128-byte row in the 128B-swizzle pattern (conflict-free: 8 consecutive rows hit 8 different 16-byte
// bank groups), fence to the async proxy, TMA-store the slab.  Rows/columns beyond the matrix are clipped by TMA.
// Returns the number of bulk groups committed (slabs inside the matrix).  `slab_ctr` carries the buffer parity.
NXD_DEVICE int epilogue_tile_tma(uint32_t tmem_acc, uint32_t stage_smem, const CUtensorMap* omap, int row0, int col0, int N,
                                 int q, int lane, bool issuer, uint32_t& slab_ctr) {
  const int row = q * 32 + lane;
  int groups = 0;
#pragma unroll 1
  for (int s = 0; s < TILE_N / kEpiSlabCols; ++s) {
    const int c0 = col0 + s * kEpiSlabCols;
    if (c0 >= N) break;
    const uint32_t buf = stage_smem + (slab_ctr & 1u) * kEpiSlabBytes;
    if (issuer) bulk_wait_read<1>();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    uint32_t r0[32], r1[32];
    tcgen05_ld_32x32(tmem_acc + s * kEpiSlabCols, r0);
    tcgen05_ld_32x32(tmem_acc + s * kEpiSlabCols + 32, r1);
    tcgen05_wait_ld();
    const uint32_t rbase = buf + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t* src = (c < 4) ? (r0 + 8 * c) : (r1 + 8 * (c - 4));
      uint4 pk;
      __nv_bfloat162* h = (__nv_bfloat162*)&pk;
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(__uint_as_float(src[2 * j]), __uint_as_float(src[2 * j + 1]));
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                   ::"r"(rbase + (((uint32_t)c ^ sw) << 4)), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (issuer) { tma_store_2d(omap, buf, c0, row0); bulk_commit(); }
    ++slab_ctr;
    ++groups;
  }
  return groups;
}

NXD_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
NXD_DEVICE void red_add_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
NXD_DEVICE uint32_t atom_add_acqrel_gpu(uint32_t* p, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}

}  // namespace g2
}  // namespace nxd
