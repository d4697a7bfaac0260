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

#include <cstdlib>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"
#include "gemm2cta.cuh"

namespace nxd {
namespace g2 {

template <bool A_KMAJOR, bool B_KMAJOR, typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      const __grid_constant__ CUtensorMap tma_out, OutT* __restrict__ out, int M, int N, int K, int accumulate,
                      int split_k) {
  // accumulate: 0 store, 1 read-modify-write, 2 atomic add (split-K: `split_k` CTA pairs share one output tile, each
  // reduces a K range and adds its fp32 partial with red.global.add.v4 — used for the weight-gradient GEMMs whose output
  // has too few tiles to fill the 74 CTA pairs while K = tokens is huge)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kBarOffset);
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
  const int kb_per = (num_kb + split_k - 1) / split_k;       // k-blocks per split (host guarantees no empty split)
  const int num_units = num_tiles * split_k;
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
      for (int unit = pair; unit < num_units; unit += num_pairs) {
        const int tile = unit % num_tiles, ks = unit / num_tiles;
        int m_blk, n_blk;
        tile_coords(tile, tiles_m, tiles_n, m_blk, n_blk);
        const int m0 = m_blk * TILE_M + (int)cta * CTA_M;
        const int n0 = n_blk * TILE_N + (int)cta * HALF_N;
        const int kb1 = min(num_kb, (ks + 1) * kb_per);
        for (int kb = ks * kb_per; kb < kb1; ++kb) {
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
      for (int unit = pair; unit < num_units; unit += num_pairs) {
        const int ks = unit / num_tiles;
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * TILE_N;
        const int kb0 = ks * kb_per, kb1 = min(num_kb, (ks + 1) * kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t da = A_KMAJOR ? make_smem_desc(sa + k * 32, 16, 1024) : make_smem_desc(sa + k * 2048, 8192, 1024);
            const uint64_t db = B_KMAJOR ? make_smem_desc(sb + k * 32, 16, 1024) : make_smem_desc(sb + k * 2048, 8192, 1024);
            tcgen05_mma_f16_2cta(tmem_d, da, db, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
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
    const bool use_tma_store = sizeof(OutT) == 2 && !accumulate;     // fp32 / read-modify-write outputs keep direct stores
    const bool issuer = threadIdx.x == 64;
    uint32_t slab_ctr = 0;
    int as = 0; uint32_t aphase = 0;
    for (int unit = pair; unit < num_units; unit += num_pairs) {
      const int tile = unit % num_tiles;
      int m_blk, n_blk;
      tile_coords(tile, tiles_m, tiles_n, m_blk, n_blk);
      mbar_wait(bar_tfull + 8 * as, aphase);
      tcgen05_fence_after();
      const int n0 = n_blk * TILE_N;
      if (use_tma_store) {
        epilogue_tile_tma(tmem_base + ((uint32_t)(q * 32) << 16) + as * TILE_N, smem_base + kEpiOffset, &tma_out,
                          m_blk * TILE_M + (int)cta * CTA_M, n0, N, q, lane, issuer, slab_ctr);
      } else {
        const int row = m_blk * TILE_M + (int)cta * CTA_M + q * 32 + lane;
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
      }
      tcgen05_fence_before();
      mbar_arrive_cluster(tempty_leader + 8 * as);   // leader's barrier collects both CTAs' epilogues
      if (++as == kAcc) { as = 0; aphase ^= 1; }
    }
    if (issuer) bulk_wait<0>();      // staging smem must outlive the last stores' reads; output complete at kernel end
  }

  tcgen05_fence_before();
  cluster_sync_all();               // peer must be done with our smem/TMEM before teardown
  if (warp == 1) { __syncwarp(); tcgen05_dealloc_2cta(tmem_base, kTmemCols); }
}


// =====================================================================================================
// Tensor-parallel fused variants on the CTA-pair kernel (no NCCL call on these paths).
//   MODE 1  all-gather → GEMM   : the first `comm_ctas` CTAs push this rank's A shard to every peer's symmetric buffer
//                                 (coalesced 16 B × 192-thread stores over NVLink) and publish per-(source,128-row block)
//                                 flags (st.release.sys = epoch); GEMM pairs walk the row chunks in arrival order, each
//                                 CTA's TMA producer acquiring the flag of its own 128 rows first.
//   MODE 2  GEMM → reduce-scatter: GEMM pairs write bf16 partial tiles to a LOCAL buffer (L2-merged stores) and bump a
//                                 local per-row-block counter; comm CTAs wait for a row block to be complete and push it to
//                                 the owner's staging slot as full 512 B rows (NVLink-friendly bursts, decoupled from the
//                                 MMA epilogue), then publish the owner's flag; finally every CTA becomes a reducer that
//                                 sums the world partials of its rank's rows in fp32.
// =====================================================================================================
struct TpComm {
  int mode, rank, world;
  const int64_t* peer_bufs;
  const int64_t* peer_flags;
  long buf_offset;
  int flag_offset;
  uint32_t epoch;
  int comm_ctas;
  int rows_per_rank;
  const void* a_local;
  void* rs_out;
  uint32_t* tile_done;
  uint32_t* gemm_done;
  uint32_t gemm_done_target;
  int push_tma;            // 0: register-staged 16-byte copies; 1: TMA bulk copies through smem, per item; 2: streaming across items
};
constexpr int kMaxRowBlocks = 64;


template <int MODE>
NXD_DEVICE void tp_tile_coords(int tile, int tiles_n, const TpComm& c, int& m_blk, int& n_blk) {
  const int mb_per_rank = c.rows_per_rank / TILE_M;           // 256-row blocks per rank
  const int per_chunk = mb_per_rank * tiles_n;
  const int step = tile / per_chunk;
  const int in = tile - step * per_chunk;
  const int chunk = MODE == 1 ? (c.rank - step + c.world) % c.world : (c.rank + 1 + step) % c.world;
  m_blk = chunk * mb_per_rank + in % mb_per_rank;
  n_blk = in / mb_per_rank;
}

// copy `bytes` (multiple of 16) with all threads of the CTA.  16 independent 16-byte loads are in flight per thread
// (48 KB per CTA) before the first store issues: NVLink/L2 latency is ~1-2 us, so bytes-in-flight is what sets the
// per-CTA copy bandwidth.
NXD_DEVICE void cta_copy16(uint8_t* dst, const uint8_t* src, size_t bytes) {
  constexpr int U = 16;
  const size_t nvec = bytes / 16;
  const uint4* s4 = (const uint4*)src;
  uint4* d4 = (uint4*)dst;
  size_t i = threadIdx.x;
  for (; i + (U - 1) * kThreads < nvec; i += U * kThreads) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = s4[i + u * kThreads];
#pragma unroll
    for (int u = 0; u < U; ++u) d4[i + u * kThreads] = v[u];
  }
  for (; i < nvec; i += kThreads) d4[i] = s4[i];
}

// TMA push: one thread streams `bytes` (multiple of 16) from local global memory to a peer address through this CTA's
// (otherwise idle) GEMM smem ring: cp.async.bulk global→smem (mbarrier completion) then smem→peer-global (bulk group).
// Five 32 KB loads and one or two stores are in flight per CTA without using a single data register, which is what the
// register-staged copy (48 KB in flight, one fence per block) could not do.  `phases` carries the ring's barrier parities
// across calls.  Barriers: the comm CTA borrows bar_empty[] (count 1) of the GEMM pipeline it never runs.
NXD_DEVICE void tma_push(uint8_t* dst, const uint8_t* src, size_t bytes, uint32_t smem_base, uint32_t bar0, uint32_t& phases) {
  constexpr uint32_t CH = (uint32_t)kStageBytes;
  constexpr int NST = kStages;
  const int n = (int)((bytes + CH - 1) / CH);
  auto chunk_bytes = [&](int c) { return (uint32_t)min((size_t)CH, bytes - (size_t)c * CH); };
  auto load = [&](int c) {
    const int st = c % NST;
    const uint32_t b = bar0 + 8 * st, nb = chunk_bytes(c);
    mbar_expect_tx(b, nb);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_base + st * CH), "l"(src + (size_t)c * CH), "r"(nb), "r"(b) : "memory");
  };
  for (int c = 0; c < n && c < NST; ++c) load(c);
  for (int c = 0; c < n; ++c) {
    const int st = c % NST;
    mbar_wait(bar0 + 8 * st, (phases >> st) & 1u);
    phases ^= 1u << st;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst + (size_t)c * CH), "r"(smem_base + st * CH), "r"(chunk_bytes(c)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (c >= 1 && c - 1 + NST < n) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // store c-1 has drained its stage
      load(c - 1 + NST);
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Streaming variant of tma_push (push_tma == 2, opt-in until measured): the smem ring is kept full ACROSS the CTA's items — no
// drain between 1 MB blocks.  Item k's flag is published as soon as all of its store groups have COMPLETED, which
// `cp.async.bulk.wait_group NST` guarantees for every group older than the newest NST ones; only when the next item is not
// ready yet (GEMM→RS) or at the very end is the pipeline drained.  `item(k, src, dst, flag)` yields the k-th item of this CTA,
// `ready(k)` blocks until its source is complete (no-op for the all-gather).
template <typename ItemFn, typename ReadyFn, typename PendingFn>
NXD_DEVICE void tma_push_stream(int n_items, size_t item_bytes, uint32_t epoch, uint32_t smem_base, uint32_t bar0,
                                uint32_t& phases, ItemFn item, ReadyFn ready, PendingFn is_ready) {
  constexpr uint32_t CH = (uint32_t)kStageBytes;
  constexpr int NST = kStages;
  const int cpi = (int)((item_bytes + CH - 1) / CH);          // chunks per item
  const long total = (long)n_items * cpi;
  int next_flag = 0;                                          // items [0, next_flag) have been published
  auto publish_upto = [&](int k_end) {                        // caller guarantees completion of those items' stores
    if (next_flag >= k_end) return;
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __threadfence_system();
    for (; next_flag < k_end; ++next_flag) {
      const uint8_t* s_; uint8_t* d_; uint32_t* f_;
      item(next_flag, s_, d_, f_);
      st_release_sys(f_, epoch);
    }
  };
  auto chunk_bytes = [&](int c) { return (uint32_t)min((size_t)CH, item_bytes - (size_t)c * CH); };
  long issued = 0;                                            // loads issued so far (global chunk index)
  int ready_upto = 0;                                         // items [0, ready_upto) are known to be loadable
  auto load = [&](long g) {
    const int k = (int)(g / cpi), c = (int)(g % cpi);
    const uint8_t* s_; uint8_t* d_; uint32_t* f_;
    item(k, s_, d_, f_);
    const int st = (int)(g % NST);
    const uint32_t b = bar0 + 8 * st, nb = chunk_bytes(c);
    mbar_expect_tx(b, nb);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_base + st * CH), "l"(s_ + (size_t)c * CH), "r"(nb), "r"(b) : "memory");
  };
  for (long g = 0; g < total; ++g) {
    // keep up to NST-1 loads ahead of the store cursor; never load from an item that is not ready
    while (issued < total && issued < g + NST - 1 + (g == 0 ? 1 : 0)) {
      const int k = (int)(issued / cpi);
      if (k >= ready_upto) {
        if (!is_ready(k)) {
          if (issued > g) break;                              // there is still loaded work to store: come back later
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // nothing in flight to overlap with: drain, publish, block
          publish_upto((int)(g / cpi));
          ready(k);
        }
        asm volatile("fence.proxy.async.global;" ::: "memory");       // generic-proxy producer writes → async-proxy reads
        ready_upto = k + 1;
      }
      if (issued >= NST) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");   // stage reuse: its old store has drained
      load(issued++);
    }
    const int k = (int)(g / cpi), c = (int)(g % cpi), st = (int)(g % NST);
    const uint8_t* s_; uint8_t* d_; uint32_t* f_;
    item(k, s_, d_, f_);
    mbar_wait(bar0 + 8 * st, (phases >> st) & 1u);
    phases ^= 1u << st;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(d_ + (size_t)c * CH), "r"(smem_base + st * CH), "r"(chunk_bytes(c)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // every group older than the newest NST has fully completed after this wait → items that ended there can be flagged
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(NST) : "memory");
    const long done_upto = g - NST;                           // global chunk index whose store is certainly complete
    if (done_upto >= 0) publish_upto((int)((done_upto + 1) / cpi));
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  publish_upto(n_items);
}

template <bool A_KMAJOR, bool B_KMAJOR, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_2cta_tp_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_a_local, __nv_bfloat16* __restrict__ out, int M, int N,
                         int K, TpComm comm) {
  using OutT = __nv_bfloat16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + kBarOffset);
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
  const bool is_comm = (int)blockIdx.x < comm.comm_ctas;
  const int pair = ((int)blockIdx.x - comm.comm_ctas) >> 1, num_pairs = ((int)gridDim.x - comm.comm_ctas) >> 1;
  const int blk128_per_rank = comm.rows_per_rank / CTA_M;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tma_a);
    prefetch_tmap(&tma_b);
    prefetch_tmap(&tma_a_local);
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
    uint32_t push_phases = 0;
    if constexpr (MODE == 1) {
      // ---- all-gather pusher: (destination step, 128-row block) items.  The local chunk is NOT copied: GEMM tiles of
      //      the own chunk read the shard in place through `tma_a_local` and start immediately. -------------------
      const int items = (comm.world - 1) * blk128_per_rank;
      const size_t blk_bytes = (size_t)CTA_M * K * 2;
      if (comm.push_tma == 2) {
        if (threadIdx.x == 0) {
          const int mine = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / comm.comm_ctas + 1 : 0;
          auto item = [&](int k, const uint8_t*& src, uint8_t*& d, uint32_t*& f) {
            const int it = (int)blockIdx.x + k * comm.comm_ctas;
            const int step = it / blk128_per_rank + 1, mb = it % blk128_per_rank;
            const int dst = (comm.rank + step) % comm.world;
            src = (const uint8_t*)comm.a_local + (size_t)mb * blk_bytes;
            d = (uint8_t*)comm.peer_bufs[dst] + comm.buf_offset + ((size_t)comm.rank * blk128_per_rank + mb) * blk_bytes;
            f = (uint32_t*)comm.peer_flags[dst] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb;
          };
          tma_push_stream(mine, blk_bytes, comm.epoch, smem_base, bar_empty, push_phases, item, [](int) {}, [](int) { return true; });
        }
      } else
      for (int it = blockIdx.x; it < items; it += comm.comm_ctas) {
        const int step = it / blk128_per_rank + 1, mb = it % blk128_per_rank;
        const int dst = (comm.rank + step) % comm.world;
        const uint8_t* src = (const uint8_t*)comm.a_local + (size_t)mb * blk_bytes;
        uint8_t* d = (uint8_t*)comm.peer_bufs[dst] + comm.buf_offset +
                     ((size_t)comm.rank * blk128_per_rank + mb) * blk_bytes;
        if (comm.push_tma) {
          if (threadIdx.x == 0) {
            tma_push(d, src, blk_bytes, smem_base, bar_empty, push_phases);
            __threadfence_system();
            st_release_sys((uint32_t*)comm.peer_flags[dst] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb, comm.epoch);
          }
        } else {
          cta_copy16(d, src, blk_bytes);
          __threadfence_system();
          __syncthreads();
          if (threadIdx.x == 0)
            st_release_sys((uint32_t*)comm.peer_flags[dst] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb, comm.epoch);
        }
      }
    } else {
      // ---- reduce-scatter pusher: ship finished row blocks of remote chunks to their owners -------------------
      const int items = (comm.world - 1) * blk128_per_rank;
      const size_t blk_bytes = (size_t)CTA_M * N * 2;
      if (comm.push_tma == 2) {
        if (threadIdx.x == 0) {
          const int mine = ((int)blockIdx.x < items) ? (items - 1 - (int)blockIdx.x) / comm.comm_ctas + 1 : 0;
          auto gblk_of = [&](int k, int& owner, int& mb) {
            const int it = (int)blockIdx.x + k * comm.comm_ctas;
            const int step = it / blk128_per_rank;
            mb = it % blk128_per_rank;
            owner = (comm.rank + 1 + step) % comm.world;
            return owner * blk128_per_rank + mb;
          };
          auto item = [&](int k, const uint8_t*& src, uint8_t*& d, uint32_t*& f) {
            int owner, mb;
            const int gblk = gblk_of(k, owner, mb);
            src = (const uint8_t*)out + (size_t)gblk * blk_bytes;
            d = (uint8_t*)comm.peer_bufs[owner] + comm.buf_offset + ((size_t)comm.rank * blk128_per_rank + mb) * blk_bytes;
            f = (uint32_t*)comm.peer_flags[owner] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb;
          };
          auto is_ready = [&](int k) {
            int owner, mb;
            const int gblk = gblk_of(k, owner, mb);
            if (ld_acquire_gpu(comm.tile_done + gblk) < (uint32_t)tiles_n) return false;
            comm.tile_done[gblk] = 0;                                  // single consumer: re-arm for the next call
            return true;
          };
          auto ready = [&](int k) { while (!is_ready(k)) __nanosleep(128); };
          tma_push_stream(mine, blk_bytes, comm.epoch, smem_base, bar_empty, push_phases, item, ready, is_ready);
        }
      } else
      for (int it = blockIdx.x; it < items; it += comm.comm_ctas) {
        const int step = it / blk128_per_rank, mb = it % blk128_per_rank;
        const int owner = (comm.rank + 1 + step) % comm.world;
        const int gblk = owner * blk128_per_rank + mb;                 // global 128-row block of the partial
        if (threadIdx.x == 0) {
          while (ld_acquire_gpu(comm.tile_done + gblk) < (uint32_t)tiles_n) __nanosleep(128);
          comm.tile_done[gblk] = 0;                                    // single consumer: re-arm for the next call
        }
        const uint8_t* src = (const uint8_t*)out + (size_t)gblk * blk_bytes;
        uint8_t* d = (uint8_t*)comm.peer_bufs[owner] + comm.buf_offset +
                     ((size_t)comm.rank * blk128_per_rank + mb) * blk_bytes;
        if (comm.push_tma) {
          if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.global;" ::: "memory");    // epilogue (generic) writes → TMA (async proxy) reads
            tma_push(d, src, blk_bytes, smem_base, bar_empty, push_phases);
            __threadfence_system();
            st_release_sys((uint32_t*)comm.peer_flags[owner] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb, comm.epoch);
          }
        } else {
          __syncthreads();
          cta_copy16(d, src, blk_bytes);
          __threadfence_system();
          __syncthreads();
          if (threadIdx.x == 0)
            st_release_sys((uint32_t*)comm.peer_flags[owner] + comm.flag_offset + comm.rank * kMaxRowBlocks + mb, comm.epoch);
        }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        tp_tile_coords<MODE>(tile, tiles_n, comm, m_blk, n_blk);
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
            const uint32_t* f = (const uint32_t*)comm.peer_flags[comm.rank] + comm.flag_offset +
                                src_rank * kMaxRowBlocks + (gblk % blk128_per_rank);
            wait_flag_ge(f, comm.epoch);
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
          if constexpr (A_KMAJOR) {
            tma_load_2d_2cta(sa, amap, full, k0, a_row);
          } else {
#pragma unroll
            for (int j = 0; j < CTA_M / 64; ++j) tma_load_2d_2cta(sa + j * 8192, amap, full, a_row + j * 64, k0);
          }
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
          tcgen05_commit_2cta(bar_empty + 8 * stage, 0b11);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_2cta(bar_tfull + 8 * as, 0b11);
        if (++as == kAcc) { as = 0; aphase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const uint32_t tempty_leader = mapa_shared(bar_tempty, 0);
    int as = 0; uint32_t aphase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tp_tile_coords<MODE>(tile, tiles_n, comm, m_blk, n_blk);
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
        if (row_ok && col0 < N) store_chunk<OutT>(orow, col0, N, r, 0);
      }
      tcgen05_fence_before();
      mbar_arrive_cluster(tempty_leader + 8 * as);
      if constexpr (MODE == 2) {
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) red_add_release_gpu(comm.tile_done + (m_blk * TILE_M + (int)cta * CTA_M) / CTA_M, 1u);
      }
      if (++as == kAcc) { as = 0; aphase ^= 1; }
    }
    if constexpr (MODE == 2) {
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) red_add_release_gpu(comm.gemm_done, 1u);
    }
  }

  if constexpr (MODE == 2) {
    // ---- reducer: (own 128-row block, 256-column slab) items, all CTAs -------------------------------------------
    __syncthreads();
    const int slabs = (N + 255) / 256;
    const int items = blk128_per_rank * slabs;
    const uint32_t* myflags = (const uint32_t*)comm.peer_flags[comm.rank] + comm.flag_offset;
    const __nv_bfloat16* staging = (const __nv_bfloat16*)((const uint8_t*)comm.peer_bufs[comm.rank] + comm.buf_offset);
    const __nv_bfloat16* own = out + (size_t)comm.rank * comm.rows_per_rank * N;
    __nv_bfloat16* rout = (__nv_bfloat16*)comm.rs_out;
    bool own_ready = false;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int mb = it / slabs, slab = it % slabs;
      if ((int)threadIdx.x < comm.world && (int)threadIdx.x != comm.rank)
        wait_flag_ge(myflags + threadIdx.x * kMaxRowBlocks + mb, comm.epoch);
      if (!own_ready && threadIdx.x == 0)
        while ((int32_t)(ld_acquire_gpu(comm.gemm_done) - comm.gemm_done_target) < 0) __nanosleep(128);
      own_ready = true;
      __syncthreads();
      const int c0 = slab * 256;
      const int cols = min(256, N - c0);
      const int vec_per_row = cols / 8;
      // 4 vectors per thread per iteration and two sources at a time → 8 independent 16-byte loads in flight
      constexpr int V = 4;
      const int total_vec = CTA_M * vec_per_row;
      const size_t src_stride = (size_t)comm.rows_per_rank * N;
      for (int base = threadIdx.x; base < total_vec; base += V * kThreads) {
        size_t off[V];
        bool ok[V];
        float acc[V][8];
#pragma unroll
        for (int u = 0; u < V; ++u) {
          const int idx = base + u * kThreads;
          ok[u] = idx < total_vec;
          const int r = ok[u] ? idx / vec_per_row : 0, v = ok[u] ? idx % vec_per_row : 0;
          off[u] = ((size_t)(mb * CTA_M + r)) * N + c0 + v * 8;
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
          const uint4 raw = ok[u] ? *(const uint4*)(own + off[u]) : make_uint4(0, 0, 0, 0);
          const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[u][2 * j] = f.x; acc[u][2 * j + 1] = f.y; }
        }
        for (int s0 = 0; s0 < comm.world; s0 += 2) {
          uint4 raw[2][V];
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const int sidx = s0 + d;
            const bool use = sidx < comm.world && sidx != comm.rank;
#pragma unroll
            for (int u = 0; u < V; ++u)
              raw[d][u] = (use && ok[u]) ? *(const uint4*)(staging + (size_t)sidx * src_stride + off[u]) : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int d = 0; d < 2; ++d)
#pragma unroll
            for (int u = 0; u < V; ++u) {
              const __nv_bfloat162* h = (const __nv_bfloat162*)&raw[d][u];
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[u][2 * j] += f.x; acc[u][2 * j + 1] += f.y; }
            }
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
          if (!ok[u]) continue;
          uint4 o; __nv_bfloat162* oh = (__nv_bfloat162*)&o;
#pragma unroll
          for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[u][2 * j], acc[u][2 * j + 1]);
          *(uint4*)(rout + off[u]) = o;
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

// ------------------------------------------------------------------ host side
CUtensorMap make_tmap_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_cols, uint32_t box_rows);
int device_sm_count();

template <bool AK, bool BK_, typename OutT>
static void launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, void* out, int M, int N, int K,
                    int accumulate, int split_k, int grid, cudaStream_t st) {
  auto kern = g2::gemm_bf16_2cta_kernel<AK, BK_, OutT>;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::kSmem));
    configured = true;
  }
  kern<<<grid, g2::kThreads, g2::kSmem, st>>>(ta, tb, to, (OutT*)out, M, N, K, accumulate, split_k);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void gemm_bf16_2cta(const void* a, const void* b, void* out, int M, int N, int K, bool trans_a, bool trans_b, int out_dt,
                    bool accumulate, cudaStream_t st) {
  const bool AK = !trans_a, BK_ = trans_b;
  const CUtensorMap ta = AK ? make_tmap_bf16(a, M, K, g2::BK, g2::CTA_M) : make_tmap_bf16(a, K, M, 64, g2::BK);
  const CUtensorMap tb = BK_ ? make_tmap_bf16(b, N, K, g2::BK, g2::HALF_N) : make_tmap_bf16(b, K, N, 64, g2::BK);
  // output map for the TMA-store epilogue (bf16, non-accumulating; needs a 16-byte row pitch) — otherwise a dummy
  const bool tma_out_ok = out_dt == kBF16 && !accumulate && N % 8 == 0;
  const CUtensorMap to = tma_out_ok ? make_tmap_bf16(out, M, N, g2::kEpiSlabCols, g2::CTA_M) : ta;
  if (out_dt == kBF16 && !accumulate && !tma_out_ok) nxd_throw("gemm_bf16_2cta: bf16 output needs N % 8 == 0", __FILE__, __LINE__);
  const int tiles = ((M + g2::TILE_M - 1) / g2::TILE_M) * ((N + g2::TILE_N - 1) / g2::TILE_N);
  const int pairs = device_sm_count() / 2;
  // Split-K for fp32 outputs (weight gradients): K = tokens is long while the output often has fewer tiles than the
  // chip has CTA pairs (TP=8: 32-176 tiles on 74 pairs).  Pick the split that minimises waves/split with a small cost per
  // extra partial (atomic epilogue traffic); every split keeps >= 16 k-blocks so the mainloop still amortises its epilogue.
  int split = 1;
  const char* split_env = getenv("NXD_GEMM_SPLITK");            // re-read per call: A/B sweeps toggle it at run time
  const bool split_enabled = !(split_env && split_env[0] == '0');
  const int num_kb = (K + g2::BK - 1) / g2::BK;
  if (out_dt == kF32 && split_enabled && N % 4 == 0) {
    double best = (double)((tiles + pairs - 1) / pairs);
    for (int s2 = 2; s2 <= 8; ++s2) {
      if (num_kb / s2 < 16) break;
      const int per = (num_kb + s2 - 1) / s2;
      if (per * (s2 - 1) >= num_kb) continue;                       // would leave an empty split
      const double cost = (double)((tiles * s2 + pairs - 1) / pairs) / s2 * (1.0 + 0.04 * (s2 - 1));
      if (cost < best - 1e-9) { best = cost; split = s2; }
    }
  }
  int acc_mode = accumulate ? 1 : 0;
  if (split > 1) {
    if (!accumulate) NXD_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)M * N * sizeof(float), st));   // partials add into zeros
    acc_mode = 2;
  }
  const int units = tiles * split;
  const int grid = 2 * (units < pairs ? units : pairs);
#define NXD_L2(AKv, BKv)                                                                          \
  do {                                                                                            \
    if (out_dt == kBF16) launch2<AKv, BKv, __nv_bfloat16>(ta, tb, to, out, M, N, K, acc_mode, split, grid, st); \
    else launch2<AKv, BKv, float>(ta, tb, to, out, M, N, K, acc_mode, split, grid, st);          \
  } while (0)
  if (AK && BK_) NXD_L2(true, true);
  else if (AK && !BK_) NXD_L2(true, false);
  else if (!AK && !BK_) NXD_L2(false, false);
  else NXD_L2(false, true);
#undef NXD_L2
}


template <bool AK, bool BK_, int MODE>
static void launch_tp(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tal, void* out, int M, int N, int K,
                      const g2::TpComm& c, int grid, cudaStream_t st) {
  auto kern = g2::gemm_bf16_2cta_tp_kernel<AK, BK_, MODE>;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::kSmem));
    configured = true;
  }
  kern<<<grid, g2::kThreads, g2::kSmem, st>>>(ta, tb, tal, (__nv_bfloat16*)out, M, N, K, c);
  NXD_CUDA_CHECK(cudaGetLastError());
}

// mode 1: `a` = this rank's gathered buffer [M,K] inside the symmetric payload, `a_local` = shard; out = [M,N].
// mode 2: `a` = [M,K]; `partial` = local [M,N] scratch; `rs_out` = [M/world, N].
void gemm_bf16_2cta_tp(int mode, const void* a, const void* b, void* out_or_partial, void* rs_out, const void* a_local,
                       int M, int N, int K, bool trans_b, int rank, int world, const int64_t* peer_bufs,
                       const int64_t* peer_flags, long buf_offset, int flag_offset, uint32_t epoch, int comm_ctas,
                       uint32_t* tile_done, uint32_t* gemm_done, uint32_t gemm_done_target, cudaStream_t st) {
  const bool BK_ = trans_b;
  const CUtensorMap ta = make_tmap_bf16(a, M, K, g2::BK, g2::CTA_M);
  const CUtensorMap tb = BK_ ? make_tmap_bf16(b, N, K, g2::BK, g2::HALF_N) : make_tmap_bf16(b, K, N, 64, g2::BK);
  g2::TpComm c{};
  c.mode = mode; c.rank = rank; c.world = world; c.peer_bufs = peer_bufs; c.peer_flags = peer_flags;
  c.buf_offset = buf_offset; c.flag_offset = flag_offset; c.epoch = epoch; c.comm_ctas = comm_ctas;
  c.rows_per_rank = M / world; c.a_local = a_local; c.rs_out = rs_out; c.tile_done = tile_done; c.gemm_done = gemm_done;
  c.gemm_done_target = gemm_done_target;
  { const char* e = getenv("NXD_TP_PUSH");      // re-read per call (sweeps): ldst | tma (default) | stream (opt-in, see tma_push_stream)
    c.push_tma = !e ? 1 : (std::string(e) == "ldst" ? 0 : (std::string(e) == "stream" ? 2 : 1)); }
  if (M % world || c.rows_per_rank % g2::TILE_M || c.rows_per_rank / g2::CTA_M > g2::kMaxRowBlocks)
    nxd_throw("fused TP GEMM (CTA-pair) needs rows/rank to be a multiple of 256 and <= 8192", __FILE__, __LINE__);
  const int grid = (device_sm_count() / 2) * 2;
  const CUtensorMap tal = (mode == 1) ? make_tmap_bf16(a_local, c.rows_per_rank, K, g2::BK, g2::CTA_M) : ta;
  if (mode == 1) { if (BK_) launch_tp<true, true, 1>(ta, tb, tal, out_or_partial, M, N, K, c, grid, st);
                   else launch_tp<true, false, 1>(ta, tb, tal, out_or_partial, M, N, K, c, grid, st); }
  else { if (BK_) launch_tp<true, true, 2>(ta, tb, tal, out_or_partial, M, N, K, c, grid, st);
         else launch_tp<true, false, 2>(ta, tb, tal, out_or_partial, M, N, K, c, grid, st); }
}

}  // namespace nxd
