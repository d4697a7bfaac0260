// bf16 GEMM for sm_100a on the 5th-generation tensor cores, optionally fused with a tensor-parallel
// collective over NVLink peer memory.
//
//   D[M,N] (+)= op(A)[M,K] · op(B)[K,N]           fp32 accumulation in TMEM
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer   : cp.async.bulk.tensor → 128B-swizzled smem ring (4 stages × 48 KB)
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (M=128, N=256, K=16) into a
//                                double-buffered 128×256 fp32 accumulator in TMEM (2×256 columns);
//                                tcgen05.commit releases smem stages / publishes accumulators
//   warps 2-5   epilogue       : tcgen05.ld (32 lanes × 32 columns per warp) → convert → global stores
//
// Operand majors are template parameters so forward (A K-major, B K-major), dgrad (B MN-major) and wgrad
// (A and B MN-major) all run without transposition copies — only the TMA boxes and UMMA descriptors differ.
//
// Fused collective modes (no NCCL call on these paths — reference call sites layers_utils.py:56,
// layers.py:1035-1038, mappings.py:125-157, 347-352):
//   MODE 1  all-gather → GEMM : `comm_sms` CTAs push this rank's A shard into every peer's symmetric
//           buffer (st.global over NVLink, 16 B vectors) and publish per-(source,row-block) flags with
//           st.release.sys; GEMM CTAs walk row-chunks in arrival order and their TMA producer acquires
//           the flag (ld.acquire.sys + fence.proxy.async) before loading tiles of that chunk.
//   MODE 2  GEMM → reduce-scatter : the epilogue stores each partial tile straight into the owner rank's
//           staging slot (peer st.global), then bumps the owner's per-(source,row-block) counter with
//           red.release.sys; after its last tile every CTA turns into a reducer that acquires the 8
//           counters of a row-block, sums the partials in fp32 and writes the bf16 result.
//   Payload regions are double-buffered by call parity and flags are monotonic, so consecutive calls need
//   no barrier ("symmetric memory protocol").
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace nxd {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;           // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int kAccStages = 2;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;   // 16 KB
constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;   // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;   // 48 KB
constexpr int kThreads = 192;
constexpr int kEpilogueThreads = 128;
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = kAccStages * BLOCK_N;  // 512
constexpr int kMaxRowBlocks = 64;   // flag index = source_rank * kMaxRowBlocks + row_block (shape independent)

struct CommDev {
  int rank, world;
  const int64_t* peer_bufs;
  const int64_t* peer_flags;
  long buf_offset;
  int flag_offset;
  uint32_t epoch;
  int comm_sms;
  int rows_per_rank;      // M / world
  uint32_t rs_targets[kMaxRowBlocks];   // MODE 2: cumulative n-tile count expected per row-block counter
  const void* a_local;    // MODE 1: this rank's [rows_per_rank, K] shard
  void* rs_out;           // MODE 2: reduced [rows_per_rank, N] output
  // grouped (MoE blockwise) modes.  MODE 3: every `grp_block_rows`-row block of A belongs to one expert whose weight
  // matrix is `grp_b_rows` rows further down the stacked 2-D B tensor.  MODE 4 (wgrad): one output matrix per expert,
  // reduced over that expert's row blocks [seg_first_block[e], seg_first_block[e+1]).
  const float* scale_a;   // FP8: per-row (token) scales [M]
  const float* scale_b;   // FP8: per-column (output channel) scales [N]
  const int* block_expert;
  const int* seg_first_block;
  int grp_block_rows, grp_b_rows, num_groups;
  long out_group_stride;
};

// MODE 4 tile → (expert, m_blk, n_blk, first k-block, number of k-blocks)
struct GroupTile { int e, m_blk, n_blk, kb0, nkb; };
NXD_DEVICE GroupTile group_tile(int tile, int tiles_m, int tiles_n, const CommDev& c) {
  GroupTile g;
  const int per = tiles_m * tiles_n;
  g.e = tile / per;
  const int in = tile - g.e * per;
  g.m_blk = in % tiles_m;
  g.n_blk = in / tiles_m;
  const int b0 = c.seg_first_block[g.e], b1 = c.seg_first_block[g.e + 1];
  const int per_blk = c.grp_block_rows / 64;
  g.kb0 = b0 * per_blk;
  g.nkb = (b1 - b0) * per_blk;
  return g;
}

// tile index → (m_blk, n_blk).  `chunk_order` visits row-chunks (one per rank) in communication order.
template <int MODE>
NXD_DEVICE void tile_coords(int tile, int tiles_m, int tiles_n, const CommDev& c, int& m_blk, int& n_blk) {
  if constexpr (MODE == 0 || MODE == 3) {
    constexpr int GROUP = 8;  // rasterise 8 M-blocks at a time so B tiles are re-used from L2
    const int per_group = GROUP * tiles_n;
    const int g = tile / per_group;
    const int first_m = g * GROUP;
    const int gsz = min(GROUP, tiles_m - first_m);
    const int in = tile - g * per_group;
    m_blk = first_m + in % gsz;
    n_blk = in / gsz;
  } else {
    const int mb_per_rank = c.rows_per_rank / BLOCK_M;
    const int per_chunk = mb_per_rank * tiles_n;
    const int step = tile / per_chunk;           // 0 .. world-1
    const int in = tile - step * per_chunk;
    int chunk;
    if constexpr (MODE == 1) chunk = (c.rank - step + c.world) % c.world;        // own, then rank-1, rank-2, …
    else chunk = (c.rank + 1 + step) % c.world;                                  // rank+1, …, own last
    m_blk = chunk * mb_per_rank + in % mb_per_rank;
    n_blk = in / mb_per_rank;
  }
}

// FP8: operands are e4m3 bytes (both K-major), one k-block is 128 elements (= the same 128-byte swizzle row), the MMA is
// kind::f8f6f4 with K = 32 per instruction, and the epilogue applies the outer product of per-row / per-column scales.
template <bool A_KMAJOR, bool B_KMAJOR, int MODE, typename OutT, bool FP8 = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 OutT* __restrict__ out, int M, int N, int K, int accumulate, CommDev comm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1 KB alignment
  uint64_t* bars = (uint64_t*)(smem + kStages * kStageBytes);
  // [0,kStages) full, [kStages,2kStages) empty, then tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 2 * kAccStages);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + kAccStages);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (M + BLOCK_M - 1) / BLOCK_M, tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = (MODE == 4 ? comm.num_groups : 1) * tiles_m * tiles_n;
  constexpr int KB_ELEMS = FP8 ? 2 * BLOCK_K : BLOCK_K;
  const int num_kb_all = (K + KB_ELEMS - 1) / KB_ELEMS;

  // ---- one-time setup ------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < kAccStages; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, kEpilogueThreads); }
    fence_barrier_init();
  }
  if (warp == 1) { tcgen05_alloc(smem_u32(tmem_slot), kTmemCols); tcgen05_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- role split ------------------------------------------------------------------
  int gemm_cta = blockIdx.x, gemm_ctas = gridDim.x;
  bool is_comm_cta = false;
  if constexpr (MODE == 1) {
    is_comm_cta = (int)blockIdx.x < comm.comm_sms;
    gemm_cta = blockIdx.x - comm.comm_sms;
    gemm_ctas = gridDim.x - comm.comm_sms;
  }

  if constexpr (MODE == 1) {
    if (is_comm_cta) {
      // ===== all-gather pusher: (destination, row-block) items, destinations in rotated order =====
      const int mb_per_rank = comm.rows_per_rank / BLOCK_M;
      const int items = comm.world * mb_per_rank;
      const size_t row_bytes = (size_t)K * 2;
      const size_t blk_bytes = (size_t)BLOCK_M * row_bytes;
      for (int it = blockIdx.x; it < items; it += comm.comm_sms) {
        const int step = it / mb_per_rank, mb = it % mb_per_rank;
        const int dst = (comm.rank + step) % comm.world;          // step 0 = local copy
        const uint8_t* src = (const uint8_t*)comm.a_local + (size_t)mb * blk_bytes;
        uint8_t* dbase = (uint8_t*)comm.peer_bufs[dst] + comm.buf_offset +
                         ((size_t)comm.rank * comm.rows_per_rank + (size_t)mb * BLOCK_M) * row_bytes;
        const size_t nvec = blk_bytes / 16;
        const uint4* s4 = (const uint4*)src;
        uint4* d4 = (uint4*)dbase;
        size_t i = threadIdx.x;
        // 4 independent 16-byte loads in flight per thread
        for (; i + 3 * kThreads < nvec; i += 4 * kThreads) {
          const uint4 v0 = __ldg(s4 + i), v1 = __ldg(s4 + i + kThreads), v2 = __ldg(s4 + i + 2 * kThreads),
                      v3 = __ldg(s4 + i + 3 * kThreads);
          d4[i] = v0; d4[i + kThreads] = v1; d4[i + 2 * kThreads] = v2; d4[i + 3 * kThreads] = v3;
        }
        for (; i < nvec; i += kThreads) d4[i] = __ldg(s4 + i);
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
          uint32_t* f = (uint32_t*)comm.peer_flags[dst] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb;
          st_release_sys(f, comm.epoch);
        }
      }
    }
  }

  if (!is_comm_cta) {
    if (warp == 0) {
      // ===== TMA producer =====
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = gemm_cta; tile < num_tiles; tile += gemm_ctas) {
          int m_blk, n_blk, kb0 = 0, num_kb = num_kb_all, b_off = 0;
          if constexpr (MODE == 4) {
            const GroupTile g = group_tile(tile, tiles_m, tiles_n, comm);
            m_blk = g.m_blk; n_blk = g.n_blk; kb0 = g.kb0; num_kb = g.nkb;
          } else {
            tile_coords<MODE>(tile, tiles_m, tiles_n, comm, m_blk, n_blk);
          }
          if constexpr (MODE == 3) b_off = comm.block_expert[(m_blk * BLOCK_M) / comm.grp_block_rows] * comm.grp_b_rows;
          if constexpr (MODE == 1) {
            const int mb_per_rank = comm.rows_per_rank / BLOCK_M;
            const uint32_t* f = (const uint32_t*)comm.peer_flags[comm.rank] + comm.flag_offset +
                                (m_blk / mb_per_rank) * kMaxRowBlocks + (m_blk % mb_per_rank);
            wait_flag_ge(f, comm.epoch);
            fence_proxy_async_global();   // peer generic-proxy writes → visible to our TMA reads
          }
          const int m0 = m_blk * BLOCK_M, n0 = n_blk * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t full = bar_full + 8 * stage;
            mbar_expect_tx(full, kStageBytes);
            const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
            const int k0 = (kb0 + kb) * KB_ELEMS;
            if constexpr (A_KMAJOR) {
              tma_load_2d(sa, &tma_a, full, k0, m0);                          // box {64 k, 128 m}
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * 8192, &tma_a, full, m0 + j * 64, k0);  // box {64 m, 64 k}
            }
            if constexpr (B_KMAJOR) {
              tma_load_2d(sb, &tma_b, full, k0, b_off + n0);                  // box {64 k, 256 n}
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * 8192, &tma_b, full, n0 + j * 64, b_off + k0);  // box {64 n, 64 k}
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer =====
      if (lane == 0) {
        constexpr uint32_t idesc = FP8 ? ((1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24))
                                       : make_idesc(!A_KMAJOR, !B_KMAJOR, BLOCK_M, BLOCK_N);
        int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
        for (int tile = gemm_cta; tile < num_tiles; tile += gemm_ctas) {
          int num_kb = num_kb_all;
          if constexpr (MODE == 4) num_kb = group_tile(tile, tiles_m, tiles_n, comm).nkb;
          mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + as * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(bar_full + 8 * stage, phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // K-major: advance 32 B inside the 128 B swizzle row; MN-major: 16 k-rows × 128 B = 2 KB
              const uint64_t da = A_KMAJOR ? make_smem_desc(sa + k * 32, 16, 1024) : make_smem_desc(sa + k * 2048, 8192, 1024);
              const uint64_t db = B_KMAJOR ? make_smem_desc(sb + k * 32, 16, 1024) : make_smem_desc(sb + k * 2048, 8192, 1024);
              if constexpr (FP8) tcgen05_mma_f8(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
              else tcgen05_mma_f16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            tcgen05_commit(bar_empty + 8 * stage);   // smem stage reusable once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          tcgen05_commit(bar_tfull + 8 * as);        // accumulator complete
          if (++as == kAccStages) { as = 0; aphase ^= 1; }
        }
      }
      __syncwarp();
    } else {
      // ===== epilogue (warps 2..5; TMEM lane quarter = warp % 4) =====
      const int q = warp & 3;
      int as = 0; uint32_t aphase = 0;
      for (int tile = gemm_cta; tile < num_tiles; tile += gemm_ctas) {
        int m_blk, n_blk;
        bool empty_group = false;
        OutT* out_base = out;
        if constexpr (MODE == 4) {
          const GroupTile g = group_tile(tile, tiles_m, tiles_n, comm);
          m_blk = g.m_blk; n_blk = g.n_blk; empty_group = g.nkb == 0;
          out_base = out + (long)g.e * comm.out_group_stride;
        } else {
          tile_coords<MODE>(tile, tiles_m, tiles_n, comm, m_blk, n_blk);
        }
        mbar_wait(bar_tfull + 8 * as, aphase);
        tcgen05_fence_after();
        const int row = m_blk * BLOCK_M + q * 32 + lane;
        const int n0 = n_blk * BLOCK_N;
        OutT* orow;
        int ld = N;
        if constexpr (MODE == 2) {
          // partial tile goes to the owner's staging slot [src = my rank][row in chunk][N]
          const int owner = (m_blk * BLOCK_M) / comm.rows_per_rank;
          const int lrow = row - owner * comm.rows_per_rank;
          uint8_t* base = (uint8_t*)comm.peer_bufs[owner] + comm.buf_offset;
          orow = (OutT*)base + ((size_t)comm.rank * comm.rows_per_rank + lrow) * (size_t)N;
        } else {
          orow = out_base + (size_t)row * ld;
        }
        const bool row_ok = row < M;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t r[32];
          tcgen05_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N + c * 32, r);
          tcgen05_wait_ld();
          if constexpr (FP8) {
            const float sa = comm.scale_a[row_ok ? row : 0];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = n0 + c * 32 + j;
              r[j] = __float_as_uint(__uint_as_float(r[j]) * sa * comm.scale_b[col < N ? col : 0]);
            }
          }
          if (MODE == 4 && empty_group) {          // expert without tokens: the accumulator was never written
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = 0u;
          }
          const int col0 = n0 + c * 32;
          if (row_ok && col0 < N) {
            if constexpr (sizeof(OutT) == 2) {
              if (col0 + 32 <= N) {
                uint4 pk[4];
                __nv_bfloat162* h = (__nv_bfloat162*)pk;
                if (accumulate && MODE != 2) {
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
                  for (int j = 0; j < 16; ++j)
                    h[j] = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                }
                uint4* dst = (uint4*)(orow + col0);
#pragma unroll
                for (int v = 0; v < 4; ++v) dst[v] = pk[v];
              } else {
                for (int j = 0; j < 32 && col0 + j < N; ++j) {
                  float v = __uint_as_float(r[j]);
                  if (accumulate && MODE != 2) v += __bfloat162float(((__nv_bfloat16*)orow)[col0 + j]);
                  ((__nv_bfloat16*)orow)[col0 + j] = __float2bfloat16_rn(v);
                }
              }
            } else {
              if (col0 + 32 <= N) {
                float4* dst = (float4*)(orow + col0);
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                  float4 o = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                         __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                  if (accumulate) { const float4 p = dst[v]; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                  dst[v] = o;
                }
              } else {
                for (int j = 0; j < 32 && col0 + j < N; ++j) {
                  float v = __uint_as_float(r[j]);
                  if (accumulate) v += ((float*)orow)[col0 + j];
                  ((float*)orow)[col0 + j] = v;
                }
              }
            }
          }
        }
        tcgen05_fence_before();
        mbar_arrive(bar_tempty + 8 * as);            // TMEM stage may be overwritten
        if constexpr (MODE == 2) {
          // publish: all 128 epilogue threads' peer stores, then one counter bump on the owner
          __threadfence_system();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (threadIdx.x == 64) {
            const int owner = (m_blk * BLOCK_M) / comm.rows_per_rank;
            const int mb_in = m_blk - owner * (comm.rows_per_rank / BLOCK_M);
            uint32_t* f = (uint32_t*)comm.peer_flags[owner] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb_in;
            red_add_release_sys(f, 1u);
          }
        }
        if (++as == kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  }

  if constexpr (MODE == 2) {
    // ===== reducer phase: (row-block, 256-column slab) items of my chunk =====
    __syncthreads();
    const int mb_per_rank = comm.rows_per_rank / BLOCK_M;
    const int slabs = (N + 255) / 256;
    const int items = mb_per_rank * slabs;
    const uint32_t* myflags = (const uint32_t*)comm.peer_flags[comm.rank] + comm.flag_offset;
    const __nv_bfloat16* stage_base = (const __nv_bfloat16*)((const uint8_t*)comm.peer_bufs[comm.rank] + comm.buf_offset);
    __nv_bfloat16* rout = (__nv_bfloat16*)comm.rs_out;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int mb = it / slabs, slab = it % slabs;
      if (threadIdx.x < comm.world) {
        // counter reaches epoch (cumulative number of n-tiles this source has ever sent for this row-block)
        wait_flag_ge(myflags + threadIdx.x * kMaxRowBlocks + mb, comm.rs_targets[mb]);
      }
      __syncthreads();
      const int c0 = slab * 256;
      const int cols = min(256, N - c0);
      const int vec_per_row = cols / 8;
      for (int idx = threadIdx.x; idx < BLOCK_M * vec_per_row; idx += kThreads) {
        const int r = idx / vec_per_row, v = idx % vec_per_row;
        const size_t off = ((size_t)(mb * BLOCK_M + r)) * N + c0 + v * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int s = 0; s < comm.world; ++s) {
          const uint4 raw = *(const uint4*)(stage_base + (size_t)s * comm.rows_per_rank * N + off);
          const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
        }
        uint4 o; __nv_bfloat162* oh = (__nv_bfloat162*)&o;
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        *(uint4*)(rout + off) = o;
      }
      __syncthreads();
    }
  }

  // ---- teardown ---------------------------------------------------------------------
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tcgen05_dealloc(tmem_base, kTmemCols); }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    NXD_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) nxd_throw("cuTensorMapEncodeTiled not available", __FILE__, __LINE__);
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// cuTensorMapEncodeTiled is a DRIVER API call: it needs a current context on the calling thread.  Autograd worker threads may
// reach us before any runtime call made the primary context current there (PyTorch's device guard skips cudaSetDevice when the
// device index already matches, and the caching allocator can serve at::empty without touching the runtime) → error 201.
static void ensure_driver_context() {
  thread_local bool done = false;
  if (!done) {
    cudaFree(nullptr);      // forces the runtime to initialise / bind the primary context on this thread
    done = true;
  }
}

// 2-D bf16 row-major tensor [rows, cols] (cols contiguous); box = {box_cols, box_rows}; 128B swizzle.
static CUtensorMap make_tmap(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  ensure_driver_context();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) nxd_throw("cuTensorMapEncodeTiled failed: " + std::to_string((int)r), __FILE__, __LINE__);
  return m;
}

// e4m3 row-major [rows, cols] bytes; box = {128 k-bytes, box_rows}
static CUtensorMap make_tmap_u8(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  ensure_driver_context();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) nxd_throw("cuTensorMapEncodeTiled (u8) failed: " + std::to_string((int)r), __FILE__, __LINE__);
  return m;
}

static int sm_count() {
  static int n = 0;
  if (!n) { int dev; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n;
}

CUtensorMap make_tmap_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows) {
  return make_tmap(ptr, rows, cols, box_cols, box_rows);
}
// same, rows `row_stride_elems` apart (strided views: fused-QKV slices, [S,B,H,D] activations)
CUtensorMap make_tmap_bf16_strided(const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                                   uint32_t box_cols, uint32_t box_rows) {
  ensure_driver_context();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) nxd_throw("cuTensorMapEncodeTiled (strided) failed: " + std::to_string((int)r), __FILE__, __LINE__);
  return m;
}
int device_sm_count() { return sm_count(); }
// packed e2m1 [rows, cols] (two elements per byte, low nibble first), loaded UNPACKED: every 16 elements (8 bytes) land in a
// 16-byte smem chunk — the operand layout kind::f8f6f4 / mxf8f6f4 expects for 4-bit types; box = {128 elements, box_rows}, so a
// k-block occupies the same 128-byte swizzle row as an fp8 one (used by gemm_mx_sm100.cu for MXFP4 weights)
CUtensorMap make_tmap_u4_unpacked_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  ensure_driver_context();
  if (cols % 128 || ((uintptr_t)ptr % 32)) nxd_throw("4-bit tensor map: K % 128 == 0 and a 32-byte aligned base", __FILE__, __LINE__);
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols / 2};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) nxd_throw("cuTensorMapEncodeTiled (u4 unpacked) failed: " + std::to_string((int)r), __FILE__, __LINE__);
  return m;
}
// packed e2m1 [rows, cols] read PACKED (two elements per byte in smem as in global memory): the operand layout of
// kind::mxf4 / kind::mxf4nvf4; box = {256 elements = 128 bytes, box_rows} (used by gemm_mxf4_sm100.cu)
CUtensorMap make_tmap_u4_packed_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  ensure_driver_context();
  if (cols % 256 || ((uintptr_t)ptr % 16)) nxd_throw("packed 4-bit tensor map: K % 256 == 0 and a 16-byte aligned base", __FILE__, __LINE__);
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols / 2};
  cuuint32_t box[2] = {256, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN8B, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) nxd_throw("cuTensorMapEncodeTiled (u4 packed) failed: " + std::to_string((int)r), __FILE__, __LINE__);
  return m;
}
// fp8 bytes, box = {128 k-bytes, box_rows} (used by gemm_mx_sm100.cu)
CUtensorMap make_tmap_u8_box(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_u8(ptr, rows, cols, box_rows);
}

template <bool AK, bool BK, int MODE, typename OutT>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, void* out, int M, int N, int K, bool accumulate,
                   const CommDev& c, int grid, cudaStream_t st) {
  auto kern = gemm_bf16_kernel<AK, BK, MODE, OutT>;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  kern<<<grid, kThreads, kSmemBytes, st>>>(ta, tb, (OutT*)out, M, N, K, accumulate ? 1 : 0, c);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void gemm_bf16(const void* a, const void* b, void* out, int M, int N, int K, bool trans_a, bool trans_b, int out_dt,
               bool accumulate, const GemmComm& comm, const void* a_local_shard, cudaStream_t st) {
  // trans_a=false: a is [M,K] (K-major).  trans_a=true: a is [K,M] (MN-major).
  // trans_b=true : b is [N,K] (K-major).  trans_b=false: b is [K,N] (MN-major).
  const bool AK = !trans_a, BK = trans_b;
  const CUtensorMap ta = AK ? make_tmap(a, M, K, BLOCK_K, BLOCK_M) : make_tmap(a, K, M, 64, BLOCK_K);
  const CUtensorMap tb = BK ? make_tmap(b, N, K, BLOCK_K, BLOCK_N) : make_tmap(b, K, N, 64, BLOCK_K);
  CommDev c{};
  c.rank = comm.rank; c.world = comm.world; c.peer_bufs = comm.peer_bufs; c.peer_flags = comm.peer_flags;
  c.buf_offset = comm.buf_offset; c.flag_offset = comm.flag_offset; c.epoch = comm.epoch; c.comm_sms = comm.comm_sms;
  c.rows_per_rank = comm.world > 0 ? M / comm.world : M;
  c.a_local = a_local_shard; c.rs_out = out;
  for (int i = 0; i < kMaxRowBlocks; ++i) c.rs_targets[i] = comm.rs_targets ? comm.rs_targets[i] : 0u;
  if (comm.mode != 0 && c.rows_per_rank / BLOCK_M > kMaxRowBlocks) nxd_throw("too many row blocks per rank", __FILE__, __LINE__);
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int sms = sm_count();
  if (comm.mode != 0) {
    if (c.rows_per_rank % BLOCK_M || M % comm.world) nxd_throw("fused TP GEMM needs rows/rank % 128 == 0", __FILE__, __LINE__);
    if (comm.mode == 2 && (N % 8)) nxd_throw("GEMM→RS needs N % 8 == 0", __FILE__, __LINE__);
  }
#define NXD_LAUNCH(AKv, BKv)                                                                                          \
  do {                                                                                                                \
    if (comm.mode == 1) {                                                                                             \
      if (out_dt == kBF16) launch<AKv, BKv, 1, __nv_bfloat16>(ta, tb, out, M, N, K, accumulate, c, sms, st);          \
      else launch<AKv, BKv, 1, float>(ta, tb, out, M, N, K, accumulate, c, sms, st);                                  \
    } else if (comm.mode == 2) {                                                                                      \
      launch<AKv, BKv, 2, __nv_bfloat16>(ta, tb, out, M, N, K, false, c, sms, st);                                    \
    } else {                                                                                                          \
      const int grid = tiles < sms ? tiles : sms;                                                                     \
      if (out_dt == kBF16) launch<AKv, BKv, 0, __nv_bfloat16>(ta, tb, out, M, N, K, accumulate, c, grid, st);         \
      else launch<AKv, BKv, 0, float>(ta, tb, out, M, N, K, accumulate, c, grid, st);                                 \
    }                                                                                                                 \
  } while (0)
  if (AK && BK) NXD_LAUNCH(true, true);
  else if (AK && !BK) NXD_LAUNCH(true, false);
  else if (!AK && !BK) NXD_LAUNCH(false, false);
  else NXD_LAUNCH(false, true);
#undef NXD_LAUNCH
}

// ---- grouped (MoE blockwise) GEMMs ---------------------------------------------------------------------------------
// out[M,N] = A[M,K] · W[e(block)]: rows of A are grouped in `block_rows`-row blocks, block i uses expert block_expert[i].
// trans_b: W is [E, N, K] (K-major) else [E, K, N] (MN-major).
void grouped_gemm_bf16(const void* a, const void* b, void* out, int M, int N, int K, int E, bool trans_b,
                       const int* block_expert, int block_rows, int out_dt, cudaStream_t st) {
  if (block_rows % BLOCK_M) nxd_throw("grouped GEMM: block size must be a multiple of 128", __FILE__, __LINE__);
  if (!trans_b && (K % BLOCK_K)) nxd_throw("grouped GEMM: K % 64 != 0 with [E,K,N] weights", __FILE__, __LINE__);
  const CUtensorMap ta = make_tmap(a, M, K, BLOCK_K, BLOCK_M);
  const CUtensorMap tb = trans_b ? make_tmap(b, (uint64_t)E * N, K, BLOCK_K, BLOCK_N) : make_tmap(b, (uint64_t)E * K, N, 64, BLOCK_K);
  CommDev c{};
  c.block_expert = block_expert; c.grp_block_rows = block_rows; c.grp_b_rows = trans_b ? N : K; c.num_groups = E;
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (trans_b) {
    if (out_dt == kBF16) launch<true, true, 3, __nv_bfloat16>(ta, tb, out, M, N, K, false, c, grid, st);
    else launch<true, true, 3, float>(ta, tb, out, M, N, K, false, c, grid, st);
  } else {
    if (out_dt == kBF16) launch<true, false, 3, __nv_bfloat16>(ta, tb, out, M, N, K, false, c, grid, st);
    else launch<true, false, 3, float>(ta, tb, out, M, N, K, false, c, grid, st);
  }
}

// out[e] [Mo, No] (+)= A[rows of e, Mo]ᵀ · B[rows of e, No]; expert e owns row blocks [seg_first_block[e], seg_first_block[e+1]).
void grouped_wgrad_bf16(const void* a, const void* b, void* out, int rows_total, int Mo, int No, int E,
                        const int* seg_first_block, int block_rows, int out_dt, bool accumulate, cudaStream_t st) {
  if (block_rows % BLOCK_K) nxd_throw("grouped wgrad: block size must be a multiple of 64", __FILE__, __LINE__);
  const CUtensorMap ta = make_tmap(a, rows_total, Mo, 64, BLOCK_K);
  const CUtensorMap tb = make_tmap(b, rows_total, No, 64, BLOCK_K);
  CommDev c{};
  c.seg_first_block = seg_first_block; c.grp_block_rows = block_rows; c.num_groups = E; c.out_group_stride = (long)Mo * No;
  const int tiles = E * ((Mo + BLOCK_M - 1) / BLOCK_M) * ((No + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (out_dt == kBF16) launch<false, false, 4, __nv_bfloat16>(ta, tb, out, Mo, No, rows_total, accumulate, c, grid, st);
  else launch<false, false, 4, float>(ta, tb, out, Mo, No, rows_total, accumulate, c, grid, st);
}

// out[M,N] (bf16) = (a_e4m3[M,K] · b_e4m3[N,K]ᵀ) ∘ scale_a[M] ⊗ scale_b[N]   (quantised inference linear)
void gemm_fp8(const void* a, const void* b, void* out, int M, int N, int K, const float* scale_a, const float* scale_b,
              cudaStream_t st) {
  if (K % 16) nxd_throw("fp8 GEMM needs K % 16 == 0", __FILE__, __LINE__);
  const CUtensorMap ta = make_tmap_u8(a, M, K, BLOCK_M);
  const CUtensorMap tb = make_tmap_u8(b, N, K, BLOCK_N);
  CommDev c{};
  c.scale_a = scale_a; c.scale_b = scale_b;
  const int tiles = ((M + BLOCK_M - 1) / BLOCK_M) * ((N + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  auto kern = gemm_bf16_kernel<true, true, 0, __nv_bfloat16, true>;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  kern<<<grid, kThreads, kSmemBytes, st>>>(ta, tb, (__nv_bfloat16*)out, M, N, K, 0, c);
  NXD_CUDA_CHECK(cudaGetLastError());
}

bool gemm_self_check_supported() { return true; }

}  // namespace nxd
