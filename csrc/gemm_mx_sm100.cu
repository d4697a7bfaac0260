// Block-scaled (OCP MX) GEMM on the 5th-generation tensor cores:  D[M,N] bf16 = (A ∘ 2^(SFA-127)) · (B ∘ 2^(SFB-127))ᵀ
//   A [M,K], B [N,K]: one fp8 byte per element (e4m3 or e5m2), K contiguous;  SFA [M,K/32], SFB [N,K/32]: E8M0 block scales.
// `tcgen05.mma.kind::mxf8f6f4.block_scale` applies the scales INSIDE the tensor core: the accumulator receives the scaled
// products, no de-quantised copy of either operand ever exists (role of the reference's MX matmul semantics,
// quantization/quantization_layers.py:626-700, experimental/quantization/microscaling/mx_torch.py:65-253).
//
// Structure = the 1-CTA bf16 kernel (gemm_sm100.cu): persistent CTAs, warp 0 TMA producer, warp 1 single-thread MMA issue,
// warps 2-5 epilogue, 6-stage smem ring, double-buffered 128×128 fp32 accumulator in TMEM.  New for block scaling:
//   * scale factors travel with their k-block: the host pre-tiles them into 512-byte chunks (128 rows × 4 blocks of 32, byte
//     offset (r%32)*16 + (r/32)*4 + j — the layout `tcgen05.cp.32x128b.warpx4` expects), one bulk copy per operand per stage;
//   * the MMA thread moves a chunk smem → TMEM with `tcgen05.cp` (4 columns per 128 rows; lane = row%32 replicated over the 4
//     sub-partitions, column = row/32, byte j = the j-th K-block of 32) right before the stage's four K=32 MMAs, which select
//     their scale byte through the a_sf_id / b_sf_id fields of the instruction descriptor;
//   * `tcgen05.cp` and `tcgen05.mma` execute in issue order, so no extra synchronisation is needed; every smem stage owns its
//     own 8 TMEM scale columns.
// Written after the round's GPU budget was spent: compiled and checked against the CUTLASS descriptor definitions
// (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled, cute/atom/mma_traits_sm100.hpp tmem_sf_frg), never executed.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace nxd {

namespace mx {

constexpr int BM = 128, BN = 128, BK = 128;              // BK elements = 128 bytes = one 128B-swizzle row
constexpr int UK = 32;                                   // K per kind::mxf8f6f4 MMA = one MX block
constexpr int kStages = 6, kAcc = 2;
constexpr int kABytes = BM * BK, kBBytes = BN * BK;      // 16 KB each
constexpr int kTileBytes = kABytes + kBBytes;            // 32 KB per stage, 1 KB aligned
constexpr int kSfBytes = 512;                            // 128 rows × 4 scale bytes
constexpr int kSfOffset = kStages * kTileBytes;          // [stage][A 512 | B 512]
constexpr int kBarOffset = kSfOffset + kStages * 2 * kSfBytes;
constexpr int kSmem = kBarOffset + 256 + 1024;
constexpr int kThreads = 192, kEpiThreads = 128;
constexpr int kTmemCols = 512;                           // 2 × 128 accumulator + 6 × 8 scale columns → next power of two
constexpr int kSfCol0 = kAcc * BN;                       // 256

NXD_DEVICE void bulk_load(uint32_t smem_dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, no swizzle: 8-row × 16-byte core matrices, rows 16 bytes apart, 8-row groups `sbo` bytes apart
NXD_DEVICE uint64_t make_smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                              // layout type 0 = SWIZZLE_NONE
}
NXD_DEVICE void tcgen05_cp_32x128b_warpx4(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
NXD_DEVICE void tcgen05_mma_mxf8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa,
                                 uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// InstrDescriptorBlockScaled: b_sf_id [4,6), a_format [7,10), b_format [10,13), a/b major 15/16 (0 = K-major), N>>3 [17,23),
// scale_format bit 23 (1 = E8M0), M>>4 [24,29), a_sf_id [29,31), k_size bit 31 (0 = K32)
__host__ __device__ constexpr uint32_t make_idesc_mx(int a_fmt, int b_fmt) {
  return ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(BN >> 3) << 17) | (1u << 23) | ((uint32_t)(BM >> 4) << 24);
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb, __nv_bfloat16* __restrict__ out, int M, int N,
                  int K, int a_fmt, int b_fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kBarOffset);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 2 * kAcc);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + kAcc);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN, num_tiles = tiles_m * tiles_n;
  const int num_kb = K / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, kEpiThreads); }
    fence_barrier_init();
  }
  if (warp == 1) { tcgen05_alloc(smem_u32(tmem_slot), kTmemCols); tcgen05_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer: operand tiles by TMA, scale chunks by bulk copy, all on the stage's full barrier =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, kTileBytes + 2 * kSfBytes);
          const uint32_t sa = smem_base + stage * kTileBytes, sb = sa + kABytes;
          const uint32_t ssf = smem_base + kSfOffset + stage * 2 * kSfBytes;
          tma_load_2d(sa, &tma_a, full, kb * BK, m_blk * BM);
          tma_load_2d(sb, &tma_b, full, kb * BK, n_blk * BN);
          bulk_load(ssf, sfa + ((size_t)m_blk * num_kb + kb) * kSfBytes, kSfBytes, full);
          bulk_load(ssf + kSfBytes, sfb + ((size_t)n_blk * num_kb + kb) * kSfBytes, kSfBytes, full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc0 = make_idesc_mx(a_fmt, b_fmt);
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait_bounded(bar_tempty + 8 * as, aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * kTileBytes, sb = sa + kABytes;
          const uint32_t ssf = smem_base + kSfOffset + stage * 2 * kSfBytes;
          const uint32_t t_sfa = tmem_base + kSfCol0 + stage * 8, t_sfb = t_sfa + 4;
          tcgen05_cp_32x128b_warpx4(t_sfa, make_smem_desc_noswizzle(ssf, 16, 128));
          tcgen05_cp_32x128b_warpx4(t_sfb, make_smem_desc_noswizzle(ssf + kSfBytes, 16, 128));
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);      // 32 bytes further inside the swizzle row
            const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
            const uint32_t idesc = idesc0 | ((uint32_t)k << 4) | ((uint32_t)k << 29);   // scale byte k of the chunk
            tcgen05_mma_mxf8(tmem_d, da, db, idesc, t_sfa, t_sfb, (kb | k) != 0 ? 1u : 0u);
          }
          tcgen05_commit(bar_empty + 8 * stage);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit(bar_tfull + 8 * as);
        if (++as == kAcc) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (warps 2..5; TMEM lane quarter = warp % 4) =====
    const int q = warp & 3;
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
      mbar_wait_bounded(bar_tfull + 8 * as, aphase);
      tcgen05_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      __nv_bfloat16* orow = out + (size_t)row * N;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tcgen05_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BN + c * 32, r);
        tcgen05_wait_ld();
        const int col0 = n_blk * BN + c * 32;
        if (row < M && col0 < N) {
          if (col0 + 32 <= N) {
            uint4 pk[4];
            __nv_bfloat162* h = (__nv_bfloat162*)pk;
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            uint4* dst = (uint4*)(orow + col0);
#pragma unroll
            for (int v = 0; v < 4; ++v) dst[v] = pk[v];
          } else {
            for (int j = 0; j < 32 && col0 + j < N; ++j) orow[col0 + j] = __float2bfloat16_rn(__uint_as_float(r[j]));
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(bar_tempty + 8 * as);
      if (++as == kAcc) { as = 0; aphase ^= 1; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tcgen05_dealloc(tmem_base, kTmemCols); }
}

}  // namespace mx

// declared in gemm_sm100.cu
CUtensorMap make_tmap_u8_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
CUtensorMap make_tmap_u4_unpacked_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
int device_sm_count();

// a [M,K], b [N,K]: fp8 bytes (format 0 = e4m3, 1 = e5m2) or packed e2m1 (format 5: K/2 bytes per row, unpacked into 8-bit
// containers by the TMA engine — W4A8: MXFP4 weights against MXFP8 activations); sfa / sfb: scale chunks tiled
// [rows/128][K/128][512]; out [M,N] bf16.  N % 8 == 0, K % 128 == 0; scale arrays padded to whole 128-row tiles by the caller.
void gemm_mxfp8(const void* a, const void* b, const void* sfa, const void* sfb, void* out, int M, int N, int K, int a_fmt,
                int b_fmt, cudaStream_t st) {
  if (K % mx::BK || N % 8) nxd_throw("gemm_mxfp8: K % 128 == 0 and N % 8 == 0", __FILE__, __LINE__);
  auto fmt_ok = [](int f) { return f == 0 || f == 1 || f == 5; };
  if (!fmt_ok(a_fmt) || !fmt_ok(b_fmt)) nxd_throw("gemm_mxfp8: formats 0 (e4m3), 1 (e5m2), 5 (e2m1)", __FILE__, __LINE__);
  const CUtensorMap ta = a_fmt == 5 ? make_tmap_u4_unpacked_box(a, M, K, mx::BM) : make_tmap_u8_box(a, M, K, mx::BM);
  const CUtensorMap tb = b_fmt == 5 ? make_tmap_u4_unpacked_box(b, N, K, mx::BN) : make_tmap_u8_box(b, N, K, mx::BN);
  const int tiles = ((M + mx::BM - 1) / mx::BM) * ((N + mx::BN - 1) / mx::BN);
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(mx::gemm_mxfp8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx::kSmem));
    configured = true;
  }
  mx::gemm_mxfp8_kernel<<<grid, mx::kThreads, mx::kSmem, st>>>(ta, tb, (const uint8_t*)sfa, (const uint8_t*)sfb,
                                                              (__nv_bfloat16*)out, M, N, K, a_fmt, b_fmt);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
