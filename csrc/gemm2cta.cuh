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
