// 4-bit block-scaled GEMM at the tensor cores' FP4 rate:  D[M,N] bf16 = alpha · (A ∘ SFA) · (B ∘ SFB)ᵀ
//   A [M,K], B [N,K]: e2m1 codes, two per byte (low nibble = even k), K contiguous — read PACKED (64 bytes per 128 elements);
//   VS = 32: OCP MXFP4 — E8M0 scale per 32 elements, `tcgen05.mma.kind::mxf4.block_scale.scale_vec::2X`
//   VS = 16: NVFP4     — UE4M3 scale per 16 elements (+ the per-tensor fp32 factor `alpha`), `kind::mxf4nvf4 … scale_vec::4X`
// Twice the math rate of kind::mxf8f6f4 (gemm_mx_sm100.cu, which unpacks e2m1 into 8-bit containers for W4A8) and half the
// operand bytes: the W4A4 path for token generation / prefill on 4-bit weights and activations (reference MX semantics:
// quantization/quantization_layers.py:626-700, experimental/quantization/microscaling/mx_torch.py:65-253).
//
// Same persistent structure as gemm_mx_sm100.cu; differences:
//   * a k-block is 256 elements = one 128-byte swizzle row of packed nibbles (TMA type 16U4_ALIGN8B), four K=64 MMAs per stage,
//     the smem descriptor advances 32 bytes per MMA exactly as for fp8;
//   * scale factors: chunks of 128 rows × 4 consecutive scales (512 bytes, the `tcgen05.cp.32x128b.warpx4` layout); a k-block
//     carries 256 / (4·VS) chunks per operand (2 for MXFP4, 4 for NVFP4), each landing in 4 TMEM columns.  An MMA consumes
//     64 / VS scales per row: MXFP4 → bytes {0,1} or {2,3} of its chunk column (a_sf_id / b_sf_id = 0 or 2), NVFP4 → a whole
//     column (sf_id 0) — cute/atom/mma_traits_sm100.hpp tmem_sf_frg, cute/arch/mma_sm100_desc.hpp (k_size 0 = K64, format E2M1 = 1).
// Written after the round's GPU budget was spent: compiled, descriptor-checked on the host, never executed.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace nxd {

namespace mxf4 {

constexpr int BM = 128, BN = 128, BK = 256;              // BK elements = 128 bytes of packed nibbles = one 128B-swizzle row
constexpr int UK = 64;                                   // K per kind::mxf4 MMA
constexpr int kStages = 6, kAcc = 2;
constexpr int kABytes = BM * BK / 2, kBBytes = BN * BK / 2;      // 16 KB each
constexpr int kTileBytes = kABytes + kBBytes;
constexpr int kSfChunk = 512;                            // 128 rows × 4 scales
constexpr int kMaxChunks = 4;                            // per operand per k-block (NVFP4); MXFP4 uses 2
constexpr int kSfStage = 2 * kMaxChunks * kSfChunk;      // 4 KB
constexpr int kSfOffset = kStages * kTileBytes;
constexpr int kBarOffset = kSfOffset + kStages * kSfStage;
constexpr int kSmem = kBarOffset + 256 + 1024;
constexpr int kThreads = 192, kEpiThreads = 128;
constexpr int kTmemCols = 512;                           // 2 × 128 accumulator + 6 × (2 operands × ≤ 16) scale columns = 448
constexpr int kSfCol0 = kAcc * BN;
constexpr int kSfColsStage = 2 * kMaxChunks * 4;         // 32

NXD_DEVICE void bulk_load(uint32_t smem_dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
NXD_DEVICE uint64_t make_smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
NXD_DEVICE void tcgen05_cp_32x128b_warpx4(uint32_t tmem_dst, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}
template <int VS>
NXD_DEVICE void tcgen05_mma_f4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa, uint32_t tmem_sfb,
                               uint32_t accumulate) {
  if constexpr (VS == 32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
  }
}
// InstrDescriptorBlockScaled for the 4-bit kinds: a_format = b_format = 1 (MXF4Format::E2M1), scale_format bit 23 (1 = UE8M0,
// 0 = UE4M3), k_size bit 31 = 0 (dense K64); N>>3 [17,23), M>>4 [24,29), b_sf_id [4,6), a_sf_id [29,31)
template <int VS>
__host__ __device__ constexpr uint32_t make_idesc_f4() {
  return (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((VS == 32 ? 1u : 0u) << 23) | ((uint32_t)(BM >> 4) << 24);
}

template <int VS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_f4_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const uint8_t* __restrict__ sfa,
               const uint8_t* __restrict__ sfb, __nv_bfloat16* __restrict__ out, int M, int N, int K, float alpha) {
  constexpr int CH = BK / (4 * VS);                      // scale chunks per operand per k-block: 2 (MXFP4) / 4 (NVFP4)
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
    // ===== producer: packed operand tiles by TMA, the k-block's CH scale chunks per operand as one bulk copy each =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % tiles_m, n_blk = tile / tiles_m;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_expect_tx(full, kTileBytes + 2 * CH * kSfChunk);
          const uint32_t sa = smem_base + stage * kTileBytes, sb = sa + kABytes;
          const uint32_t ssf = smem_base + kSfOffset + stage * kSfStage;
          tma_load_2d(sa, &tma_a, full, kb * BK, m_blk * BM);
          tma_load_2d(sb, &tma_b, full, kb * BK, n_blk * BN);
          bulk_load(ssf, sfa + ((size_t)m_blk * num_kb + kb) * (CH * kSfChunk), CH * kSfChunk, full);
          bulk_load(ssf + kMaxChunks * kSfChunk, sfb + ((size_t)n_blk * num_kb + kb) * (CH * kSfChunk), CH * kSfChunk, full);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc0 = make_idesc_f4<VS>();
      int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait_bounded(bar_tempty + 8 * as, aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * kTileBytes, sb = sa + kABytes;
          const uint32_t ssf = smem_base + kSfOffset + stage * kSfStage;
          const uint32_t t_sfa = tmem_base + kSfCol0 + stage * kSfColsStage, t_sfb = t_sfa + kMaxChunks * 4;
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            tcgen05_cp_32x128b_warpx4(t_sfa + 4 * c, make_smem_desc_noswizzle(ssf + c * kSfChunk, 16, 128));
            tcgen05_cp_32x128b_warpx4(t_sfb + 4 * c, make_smem_desc_noswizzle(ssf + (kMaxChunks + c) * kSfChunk, 16, 128));
          }
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);      // 64 nibbles = 32 bytes further inside the swizzle row
            const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
            // MXFP4: two MMAs share a chunk column (scale bytes {0,1} then {2,3}); NVFP4: one chunk column per MMA
            const int chunk = VS == 32 ? k >> 1 : k;
            const uint32_t sf_id = VS == 32 ? (uint32_t)(k & 1) * 2u : 0u;
            const uint32_t idesc = idesc0 | (sf_id << 4) | (sf_id << 29);
            tcgen05_mma_f4<VS>(tmem_d, da, db, idesc, t_sfa + 4 * chunk, t_sfb + 4 * chunk, (kb | k) != 0 ? 1u : 0u);
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
    // ===== epilogue (warps 2..5; TMEM lane quarter = warp % 4): alpha, bf16, 64-byte row segments =====
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
            for (int j = 0; j < 16; ++j)
              h[j] = __floats2bfloat162_rn(alpha * __uint_as_float(r[2 * j]), alpha * __uint_as_float(r[2 * j + 1]));
            uint4* dst = (uint4*)(orow + col0);
#pragma unroll
            for (int v = 0; v < 4; ++v) dst[v] = pk[v];
          } else {
            for (int j = 0; j < 32 && col0 + j < N; ++j) orow[col0 + j] = __float2bfloat16_rn(alpha * __uint_as_float(r[j]));
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

}  // namespace mxf4

// declared in gemm_sm100.cu
CUtensorMap make_tmap_u4_packed_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
int device_sm_count();

// a [M,K], b [N,K]: packed e2m1 (K/2 bytes per row); sfa / sfb: scale chunks tiled [rows/128][K/(4·VS)][512] (E8M0 for
// vec_size 32, UE4M3 for vec_size 16); out [M,N] bf16 = alpha · product.  K % 256 == 0, N % 8 == 0.
void gemm_f4(const void* a, const void* b, const void* sfa, const void* sfb, void* out, int M, int N, int K, int vec_size, float alpha,
             cudaStream_t st) {
  if (K % mxf4::BK || N % 8) nxd_throw("gemm_f4: K % 256 == 0 and N % 8 == 0", __FILE__, __LINE__);
  if (vec_size != 32 && vec_size != 16) nxd_throw("gemm_f4: scale vector size 32 (MXFP4) or 16 (NVFP4)", __FILE__, __LINE__);
  const CUtensorMap ta = make_tmap_u4_packed_box(a, M, K, mxf4::BM);
  const CUtensorMap tb = make_tmap_u4_packed_box(b, N, K, mxf4::BN);
  const int tiles = ((M + mxf4::BM - 1) / mxf4::BM) * ((N + mxf4::BN - 1) / mxf4::BN);
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(mxf4::gemm_f4_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mxf4::kSmem));
    NXD_CUDA_CHECK(cudaFuncSetAttribute(mxf4::gemm_f4_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, mxf4::kSmem));
    configured = true;
  }
  if (vec_size == 32)
    mxf4::gemm_f4_kernel<32><<<grid, mxf4::kThreads, mxf4::kSmem, st>>>(ta, tb, (const uint8_t*)sfa, (const uint8_t*)sfb,
                                                                         (__nv_bfloat16*)out, M, N, K, alpha);
  else
    mxf4::gemm_f4_kernel<16><<<grid, mxf4::kThreads, mxf4::kSmem, st>>>(ta, tb, (const uint8_t*)sfa, (const uint8_t*)sfb,
                                                                         (__nv_bfloat16*)out, M, N, K, alpha);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
