// Tensor-parallel GEMM fused with its collective over NVLS (NVSwitch multicast / in-switch reduction), CTA-pair
// tcgen05 kernel (cta_group::2, 256x256 tiles, TMEM double-buffered, TMA-fed — same GEMM core as gemm2cta_sm100.cu).
// No NCCL call and no staging copies on these paths:
//
//   MODE 1  all-gather → GEMM   (ColumnParallelLinear fwd with sequence parallel, RowParallelLinear dgrad;
//           reference: layers_utils.py:56 + layers.py:461, mappings.py:347-352)
//           A few "comm" CTAs stream this rank's A shard to the MULTICAST address of the symmetric buffer with
//           `multimem.st`: one store is replicated by the switch into all ranks' buffers (this rank's egress is 1x the
//           shard instead of (world-1)x), then publish per-(source, 128-row block, quarter) epoch flags with
//           st.release.sys.  GEMM pairs start on the own chunk (read in place) and then walk the remote chunks in
//           arrival order; the TMA producer acquires the four quarter flags of its 128 rows before loading them.
//
//   MODE 2  GEMM → reduce-scatter (RowParallelLinear fwd with sequence parallel, ColumnParallelLinear dgrad;
//           reference: layers.py:1031-1045, layers_utils.py:132)
//           Every rank walks the output row blocks in the SAME global order and writes its partial tiles into its own
//           symmetric buffer; the CTA that completes a 128-row block releases one flag at the block's owner.  The owner
//           pulls the sum of all ranks' partials with `multimem.ld_reduce` — the switch fetches the `world` copies and
//           adds them in fp32 (`.acc::f32`), so there is no staging buffer, no local reduction pass and no precision loss
//           from rounding between partial sums.  `wire_fp32` keeps the partials in fp32 end to end (the reference's
//           reduce_dtype=fp32 wire format, 2x the bytes).  Reduction work items are claimed dynamically: the comm CTAs
//           reduce while the GEMM runs, GEMM CTAs join when their tiles are done.
//
// Without a multicast mapping (`mc_base == nullptr`: single-GPU loopback tests, fabrics without NVLS) the same kernel
// falls back to unicast peer stores / peer loads with the identical flag protocol.
#include <cuda.h>

#include <cstdlib>
#include <string>
#include <type_traits>

#include "common.cuh"
#include "gemm2cta.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace nxd {
namespace g2 {

constexpr int kNvlsMaxRowBlocks = 256;      // 128-row blocks per rank → rows/rank <= 32768
constexpr int kAgParts = 4;                 // all-gather flag granularity: quarter of a 128-row block
constexpr int kRsRowsPerItem = 16;          // reduce work item: 16 full rows

struct NvlsComm {
  int rank, world;
  const int64_t* peer_bases;   // device table [world]: unicast VA of every rank's symmetric region
  uint8_t* mc_base;            // multicast VA of the region (nullptr → unicast fallback)
  uint8_t* local_base;         // this rank's own VA of the region
  long buf_offset;             // payload of this call inside the region (bytes)
  long flag_offset;            // flags of this op inside the region (bytes)
  uint32_t epoch;
  int comm_ctas;
  int rows_per_rank;
  const void* a_local;         // MODE 1: this rank's shard
  void* rs_out;                // MODE 2: [rows_per_rank, N] bf16
  uint32_t* tile_done;         // MODE 2: local per-128-row-block tile counters
  uint32_t* claim;             // MODE 2: reducer work counter (monotonic across calls)
  uint32_t claim_base;
  int gemm_join;               // MODE 2: GEMM CTAs become reducers after their tiles (1) or exit and free their SMs for a
                               // concurrently launched kernel, e.g. the weight-gradient GEMM on a side stream (0)
};

NXD_DEVICE void multimem_st_v4(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
NXD_DEVICE uint4 multimem_ld_reduce_bf16x2(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
NXD_DEVICE uint4 multimem_ld_reduce_f32(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
NXD_DEVICE uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
NXD_DEVICE uint4 ld_relaxed_sys_v4(const void* p) {      // peer payload published through a flag: never the nc path
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

// MODE 1: own chunk first (no wait), then the remote chunks block-major in the order their pushes were issued.
NXD_DEVICE void nvls_coords_ag(int tile, int tiles_n, const NvlsComm& c, int& m_blk, int& n_blk) {
  const int mbr = c.rows_per_rank / TILE_M;
  const int seq = tile / tiles_n;
  n_blk = tile - seq * tiles_n;
  if (seq < mbr) { m_blk = c.rank * mbr + seq; return; }
  const int s2 = seq - mbr;
  const int mb = s2 / (c.world - 1), s = s2 - mb * (c.world - 1);
  m_blk = ((c.rank + 1 + s) % c.world) * mbr + mb;
}
// MODE 2: identical order on every rank — block b of every owner completes everywhere at about the same time, so the
// owners' pulls are spread evenly over the GEMM instead of piling up at the end.
NXD_DEVICE void nvls_coords_rs(int tile, int tiles_n, const NvlsComm& c, int& m_blk, int& n_blk) {
  const int mbr = c.rows_per_rank / TILE_M;
  const int seq = tile / tiles_n;
  n_blk = tile - seq * tiles_n;
  const int mb = seq / c.world, chunk = seq - mb * c.world;
  m_blk = chunk * mbr + mb;
}

template <bool B_KMAJOR, int MODE, bool WIRE32>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_2cta_nvls_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                           const __grid_constant__ CUtensorMap tma_a_local, const __grid_constant__ CUtensorMap tma_out,
                           void* __restrict__ out_raw, int M, int N, int K, NvlsComm comm) {
  using OutT = typename std::conditional<(MODE == 2 && WIRE32), float, __nv_bfloat16>::type;
  OutT* out = (OutT*)out_raw;
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_item;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kBarOffset);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * kStages + 2 * kAcc);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStages);
  const uint32_t bar_tfull = smem_u32(bars + 2 * kStages), bar_tempty = smem_u32(bars + 2 * kStages + kAcc);
  const uint32_t smem_base = smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int tiles_m = M / TILE_M, tiles_n = (N + TILE_N - 1) / TILE_N;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (K + BK - 1) / BK;
  const bool is_comm = (int)blockIdx.x < comm.comm_ctas;
  const int pair = ((int)blockIdx.x - comm.comm_ctas) >> 1, num_pairs = ((int)gridDim.x - comm.comm_ctas) >> 1;
  const int blk128_per_rank = comm.rows_per_rank / CTA_M;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    prefetch_tmap(&tma_a_local);
    prefetch_tmap(&tma_out);
    for (int i = 0; i < kStages; ++i) { mbar_init(bar_full + 8 * i, 2); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 2 * 128); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) { tcgen05_alloc_2cta(smem_u32(tmem_slot), kTmemCols); tcgen05_relinquish_2cta(); }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (is_comm) {
    if constexpr (MODE == 1) {
      // ---- all-gather pusher -----------------------------------------------------------------------------------
      constexpr int U = 8;
      constexpr int rows_part = CTA_M / kAgParts;
      const int items = blk128_per_rank * kAgParts;
      const size_t nvec = (size_t)rows_part * K / 8;          // 16-byte vectors per item
      int ring_pos = 0;
      uint32_t ring_phases = 0;
      for (int it = blockIdx.x; it < items; it += comm.comm_ctas) {
        const int mb = it / kAgParts, part = it - mb * kAgParts;
        const size_t row0 = (size_t)mb * CTA_M + part * rows_part;
        const uint4* src = (const uint4*)((const uint8_t*)comm.a_local + row0 * K * 2);
        const size_t dst_off = (size_t)comm.buf_offset + ((size_t)comm.rank * comm.rows_per_rank + row0) * K * 2;
        if (comm.mc_base != nullptr) {
          // Stage the item through this CTA's (otherwise idle) 192 KB GEMM ring with TMA bulk loads — six 32 KB chunks
          // in flight per CTA without a data register — and fan it out with multimem.st from shared memory.  (The first
          // version kept 8 x 16 B per thread in registers: 24 KB in flight per CTA, measured 13 GB/s per CTA.)
          constexpr uint32_t CH = (uint32_t)kStageBytes;
          const uint32_t item_bytes = (uint32_t)(nvec * 16);
          const int nch = (int)((item_bytes + CH - 1) / CH);
          auto chunk_bytes = [&](int c) { return min(CH, item_bytes - (uint32_t)c * CH); };
          uint8_t* dst = comm.mc_base + dst_off;
          // prologue: fill the ring (stages are free: the previous item ended with a CTA barrier after its last reads)
          if (threadIdx.x == 0) {
            for (int c = 0; c < nch && c < kStages; ++c) {
              const uint32_t b = bar_empty + 8 * ((ring_pos + c) % kStages), nb = chunk_bytes(c);
              mbar_expect_tx(b, nb);
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(smem_base + ((ring_pos + c) % kStages) * CH), "l"((const uint8_t*)src + (size_t)c * CH), "r"(nb), "r"(b)
                           : "memory");
            }
          }
          for (int c = 0; c < nch; ++c) {
            const int st = (ring_pos + c) % kStages;
            mbar_wait(bar_empty + 8 * st, (ring_phases >> st) & 1u);
            ring_phases ^= 1u << st;
            const uint32_t nb = chunk_bytes(c);
            const uint32_t sbase = smem_base + st * CH;
            uint8_t* d = dst + (size_t)c * CH;
            for (uint32_t o = threadIdx.x * 16u; o < nb; o += kThreads * 16u) {
              uint4 v;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sbase + o));
              multimem_st_v4(d + o, v);
            }
            if (c + kStages < nch) {
              __syncthreads();                                   // everyone has read this stage: refill it
              if (threadIdx.x == 0) {
                const uint32_t b = bar_empty + 8 * st, nb2 = chunk_bytes(c + kStages);
                mbar_expect_tx(b, nb2);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sbase), "l"((const uint8_t*)src + (size_t)(c + kStages) * CH), "r"(nb2), "r"(b) : "memory");
              }
            }
          }
          ring_pos = (ring_pos + nch) % kStages;
        } else {
          for (size_t base = threadIdx.x; base < nvec; base += (size_t)U * kThreads) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const size_t i = base + (size_t)u * kThreads; if (i < nvec) v[u] = ld_nc_v4(src + i); }
            for (int p = 0; p < comm.world; ++p) {
              uint4* dst = (uint4*)((uint8_t*)comm.peer_bases[p] + dst_off);
#pragma unroll
              for (int u = 0; u < U; ++u) { const size_t i = base + (size_t)u * kThreads; if (i < nvec) dst[i] = v[u]; }
            }
          }
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < comm.world && (int)threadIdx.x != comm.rank) {
          uint32_t* f = (uint32_t*)((uint8_t*)comm.peer_bases[threadIdx.x] + comm.flag_offset) +
                        ((size_t)comm.rank * kNvlsMaxRowBlocks + mb) * kAgParts + part;
          st_release_sys(f, comm.epoch);
        }
      }
    }
    // MODE 2: comm CTAs go straight to the reducer loop below
  } else if (warp == 0) {
    // ===== TMA producer (both CTAs of the pair) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        if constexpr (MODE == 1) nvls_coords_ag(tile, tiles_n, comm, m_blk, n_blk);
        else nvls_coords_rs(tile, tiles_n, comm, m_blk, n_blk);
        const int m0 = m_blk * TILE_M + (int)cta * CTA_M;
        const int n0 = n_blk * TILE_N + (int)cta * HALF_N;
        const CUtensorMap* amap = &tma_a;
        int a_row = m0;
        if constexpr (MODE == 1) {
          const int gblk = m0 / CTA_M;
          const int src_rank = gblk / blk128_per_rank;
          if (src_rank == comm.rank) {
            amap = &tma_a_local;                       // own shard, in place
            a_row = m0 - src_rank * comm.rows_per_rank;
          } else {
            const uint32_t* f = (const uint32_t*)(comm.local_base + comm.flag_offset) +
                                ((size_t)src_rank * kNvlsMaxRowBlocks + (gblk - src_rank * blk128_per_rank)) * kAgParts;
#pragma unroll
            for (int p = 0; p < kAgParts; ++p) wait_flag_ge(f + p, comm.epoch);
            fence_proxy_async_global();
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          if (leader) mbar_expect_tx(full, 2 * kStageBytes);
          else mbar_arrive_cluster(mapa_shared(full, 0));
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
          const int k0 = kb * BK;
          tma_load_2d_2cta(sa, amap, full, k0, a_row);
          if constexpr (B_KMAJOR) {
            tma_load_2d_2cta(sb, &tma_b, full, k0, n0);
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
      constexpr uint32_t idesc = make_idesc(false, !B_KMAJOR, TILE_M, TILE_N);
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
            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_KMAJOR ? make_smem_desc(sb + k * 32, 16, 1024) : make_smem_desc(sb + k * 2048, 8192, 1024);
            tcgen05_mma_f16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tcgen05_commit_2cta(bar_empty + 8 * stage, 0b11);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_2cta(bar_tfull + 8 * as, 0b11);
        if (++as == kAcc) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (both CTAs; own 128 TMEM lanes) =====
    const int q = warp & 3;
    const uint32_t tempty_leader = mapa_shared(bar_tempty, 0);
    const bool issuer = threadIdx.x == 64;
    uint32_t slab_ctr = 0;
    int prev_gblk = -1;
    // MODE 2: a finished 128-row block of partials is announced to its owner (by whichever CTA stored its last tile)
    auto signal_block = [&](int gblk) {
      if (atom_add_acqrel_gpu(comm.tile_done + gblk, 1u) == (uint32_t)(tiles_n - 1)) {
        comm.tile_done[gblk] = 0;                                    // re-arm for the next call
        const int owner = gblk / blk128_per_rank, mb = gblk - owner * blk128_per_rank;
        uint32_t* f = (uint32_t*)((uint8_t*)comm.peer_bases[owner] + comm.flag_offset) + (size_t)comm.rank * kNvlsMaxRowBlocks + mb;
        st_release_sys(f, comm.epoch);
      }
    };
    int as = 0; uint32_t aphase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      if constexpr (MODE == 1) nvls_coords_ag(tile, tiles_n, comm, m_blk, n_blk);
      else nvls_coords_rs(tile, tiles_n, comm, m_blk, n_blk);
      mbar_wait(bar_tfull + 8 * as, aphase);
      tcgen05_fence_after();
      const int row0 = m_blk * TILE_M + (int)cta * CTA_M;
      const int n0 = n_blk * TILE_N;
      if constexpr (MODE == 2 && WIRE32) {
        OutT* orow = out + (size_t)(row0 + q * 32 + lane) * N;
#pragma unroll 1
        for (int c = 0; c < TILE_N / 32; ++c) {
          uint32_t r[32];
          tcgen05_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * TILE_N + c * 32, r);
          tcgen05_wait_ld();
          const int col0 = n0 + c * 32;
          if (col0 < N) store_chunk<OutT>(orow, col0, N, r, 0);
        }
        tcgen05_fence_before();
        mbar_arrive_cluster(tempty_leader + 8 * as);
        __threadfence_system();                                       // partial tile visible system-wide, then count it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (issuer) signal_block(row0 / CTA_M);
      } else {
        const int groups = epilogue_tile_tma(tmem_base + ((uint32_t)(q * 32) << 16) + as * TILE_N, smem_base + kEpiOffset, &tma_out,
                                             row0, n0, N, q, lane, issuer, slab_ctr);
        tcgen05_fence_before();
        mbar_arrive_cluster(tempty_leader + 8 * as);
        if constexpr (MODE == 2) {
          // the TMA stores of THIS tile stay in flight; the previous tile's are complete once at most `groups` bulk groups
          // are pending → announce the previous tile now (one tile of latency instead of a drain per tile)
          if (issuer) {
            if (prev_gblk >= 0) {
              bulk_wait_n(groups);
              fence_proxy_async_global();
              __threadfence_system();
              signal_block(prev_gblk);
            }
            prev_gblk = row0 / CTA_M;
          }
        }
      }
      if (++as == kAcc) { as = 0; aphase ^= 1; }
    }
    if (issuer) {
      bulk_wait<0>();
      if constexpr (MODE == 2 && !WIRE32) {
        if (prev_gblk >= 0) {
          fence_proxy_async_global();
          __threadfence_system();
          signal_block(prev_gblk);
        }
      }
    }
  }

  if constexpr (MODE == 2) {
    // ---- reducer: dynamically claimed (128-row block, 16-row slice) items of THIS rank's rows ------------------------
    constexpr int U = 8;
    constexpr int parts = CTA_M / kRsRowsPerItem;
    const int items = blk128_per_rank * parts;
    const uint32_t* myflags = (const uint32_t*)(comm.local_base + comm.flag_offset);
    __nv_bfloat16* rout = (__nv_bfloat16*)comm.rs_out;
    const size_t esz = WIRE32 ? 4 : 2;
    const size_t groups = (size_t)kRsRowsPerItem * N / 8;           // 8 output elements (16 bytes of bf16) per group
    __syncthreads();
    if (is_comm || comm.gemm_join)
    for (;;) {
      if (threadIdx.x == 0) s_item = (int)(atomicAdd(comm.claim, 1u) - comm.claim_base);
      __syncthreads();
      const int it = s_item;
      if (it >= items) break;
      const int mb = it / parts, part = it - mb * parts;
      if ((int)threadIdx.x < comm.world) wait_flag_ge(myflags + (size_t)threadIdx.x * kNvlsMaxRowBlocks + mb, comm.epoch);
      __syncthreads();
      const size_t row0 = (size_t)mb * CTA_M + (size_t)part * kRsRowsPerItem;
      const size_t src_off = (size_t)comm.buf_offset + ((size_t)comm.rank * comm.rows_per_rank + row0) * N * esz;
      uint4* dst = (uint4*)(rout + row0 * N);
      if (comm.mc_base != nullptr) {
        const uint8_t* src = comm.mc_base + src_off;
        for (size_t base = threadIdx.x; base < groups; base += (size_t)U * kThreads) {
          uint4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t i = base + (size_t)u * kThreads;
            if (i < groups) {
              if constexpr (WIRE32) {
                const uint4 lo = multimem_ld_reduce_f32(src + i * 32), hi = multimem_ld_reduce_f32(src + i * 32 + 16);
                __nv_bfloat162* h = (__nv_bfloat162*)&v[u];
                h[0] = __floats2bfloat162_rn(__uint_as_float(lo.x), __uint_as_float(lo.y));
                h[1] = __floats2bfloat162_rn(__uint_as_float(lo.z), __uint_as_float(lo.w));
                h[2] = __floats2bfloat162_rn(__uint_as_float(hi.x), __uint_as_float(hi.y));
                h[3] = __floats2bfloat162_rn(__uint_as_float(hi.z), __uint_as_float(hi.w));
              } else {
                v[u] = multimem_ld_reduce_bf16x2(src + i * 16);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) { const size_t i = base + (size_t)u * kThreads; if (i < groups) dst[i] = v[u]; }
        }
      } else {
        // unicast pull: read every rank's partial over NVLink (or locally) and add in fp32, fixed rank order
        for (size_t i = threadIdx.x; i < groups; i += kThreads) {
          float acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = 0.f;
          for (int p = 0; p < comm.world; ++p) {
            const uint8_t* src = (const uint8_t*)comm.peer_bases[p] + src_off;
            if constexpr (WIRE32) {
              const uint4 lo = ld_relaxed_sys_v4(src + i * 32), hi = ld_relaxed_sys_v4(src + i * 32 + 16);
              acc[0] += __uint_as_float(lo.x); acc[1] += __uint_as_float(lo.y); acc[2] += __uint_as_float(lo.z); acc[3] += __uint_as_float(lo.w);
              acc[4] += __uint_as_float(hi.x); acc[5] += __uint_as_float(hi.y); acc[6] += __uint_as_float(hi.z); acc[7] += __uint_as_float(hi.w);
            } else {
              const uint4 raw = ld_relaxed_sys_v4(src + i * 16);
              const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
            }
          }
          uint4 o; __nv_bfloat162* oh = (__nv_bfloat162*)&o;
#pragma unroll
          for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
          dst[i] = o;
        }
      }
      __syncthreads();
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) { __syncwarp(); tcgen05_dealloc_2cta(tmem_base, kTmemCols); }
}

}  // namespace g2

CUtensorMap make_tmap_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);
int device_sm_count();

template <bool BK_, int MODE, bool W32>
static void launch_nvls(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal, const CUtensorMap& to, void* out, int M,
                        int N, int K, const g2::NvlsComm& c, int grid, cudaStream_t st) {
  auto kern = g2::gemm_bf16_2cta_nvls_kernel<BK_, MODE, W32>;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::kSmem));
    configured = true;
  }
  kern<<<grid, g2::kThreads, g2::kSmem, st>>>(ta, tb, tal, to, out, M, N, K, c);
  NXD_CUDA_CHECK(cudaGetLastError());
}

// mode 1: `a_local` = shard [M/world, K]; the gathered A is the symmetric payload at buf_offset; out = [M, N] bf16.
// mode 2: `a` = [M, K]; partials go to the symmetric payload at buf_offset; rs_out = [M/world, N] bf16.
// Returns the number of reducer claims this launch consumes (mode 2; the caller advances its claim base by it).
int gemm_bf16_2cta_nvls(int mode, const void* a, const void* b, void* out, void* rs_out, const void* a_local, int M, int N, int K,
                        bool trans_b, int rank, int world, const int64_t* peer_bases, int64_t mc_base, int64_t local_base,
                        long buf_offset, long flag_offset, uint32_t epoch, int comm_ctas, uint32_t* tile_done, uint32_t* claim,
                        uint32_t claim_base, bool wire_fp32, bool gemm_join, cudaStream_t st) {
  g2::NvlsComm c{};
  c.rank = rank; c.world = world; c.peer_bases = peer_bases; c.mc_base = (uint8_t*)mc_base; c.local_base = (uint8_t*)local_base;
  c.buf_offset = buf_offset; c.flag_offset = flag_offset; c.epoch = epoch; c.comm_ctas = comm_ctas & ~1;
  c.rows_per_rank = M / world; c.a_local = a_local; c.rs_out = rs_out; c.tile_done = tile_done; c.claim = claim;
  c.claim_base = claim_base;
  c.gemm_join = (gemm_join || c.comm_ctas == 0) ? 1 : 0;
  if (M % world || c.rows_per_rank % g2::TILE_M || c.rows_per_rank / g2::CTA_M > g2::kNvlsMaxRowBlocks || N % 8 || K % 8)
    nxd_throw("fused TP GEMM (NVLS) needs rows/rank to be a multiple of 256 and <= 32768, N and K multiples of 8", __FILE__, __LINE__);
  const int grid = (device_sm_count() / 2) * 2;
  if (c.comm_ctas >= grid) nxd_throw("comm_ctas leaves no GEMM CTAs", __FILE__, __LINE__);
  if (mode == 1 && c.comm_ctas < 2) c.comm_ctas = 2;
  const uint8_t* payload = (const uint8_t*)local_base + buf_offset;
  if (mode == 1) {
    const CUtensorMap ta = make_tmap_bf16(payload, M, K, g2::BK, g2::CTA_M);
    const CUtensorMap tal = make_tmap_bf16(a_local, c.rows_per_rank, K, g2::BK, g2::CTA_M);
    const CUtensorMap tb = trans_b ? make_tmap_bf16(b, N, K, g2::BK, g2::HALF_N) : make_tmap_bf16(b, K, N, 64, g2::BK);
    const CUtensorMap to = make_tmap_bf16(out, M, N, g2::kEpiSlabCols, g2::CTA_M);
    if (trans_b) launch_nvls<true, 1, false>(ta, tb, tal, to, out, M, N, K, c, grid, st);
    else launch_nvls<false, 1, false>(ta, tb, tal, to, out, M, N, K, c, grid, st);
    return 0;
  }
  const CUtensorMap ta = make_tmap_bf16(a, M, K, g2::BK, g2::CTA_M);
  const CUtensorMap tb = trans_b ? make_tmap_bf16(b, N, K, g2::BK, g2::HALF_N) : make_tmap_bf16(b, K, N, 64, g2::BK);
  void* partial = (void*)payload;
  const CUtensorMap to = wire_fp32 ? ta : make_tmap_bf16(partial, M, N, g2::kEpiSlabCols, g2::CTA_M);
  if (trans_b) { if (wire_fp32) launch_nvls<true, 2, true>(ta, tb, ta, to, partial, M, N, K, c, grid, st);
                 else launch_nvls<true, 2, false>(ta, tb, ta, to, partial, M, N, K, c, grid, st); }
  else { if (wire_fp32) launch_nvls<false, 2, true>(ta, tb, ta, to, partial, M, N, K, c, grid, st);
         else launch_nvls<false, 2, false>(ta, tb, ta, to, partial, M, N, K, c, grid, st); }
  return (c.rows_per_rank / g2::CTA_M) * (g2::CTA_M / g2::kRsRowsPerItem) + (c.gemm_join ? grid : c.comm_ctas);
}

}  // namespace nxd
