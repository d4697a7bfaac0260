// Flash attention for sm_100a (head_dim 128, bf16): tcgen05 QKᵀ / PV with TMEM accumulators, TMA-staged operands,
// online softmax with one thread per query row.  Role parity: reference kernels/flash_attn.py:162-212
// (nki_flash_attn_func → flash_fwd / flash_attn_bwd) and modules/attention call sites.
//
// Forward, one CTA per (128-query tile, head, batch), one CTA per SM (TMEM: S double buffer + O = 384 of 512 columns):
//   warp 0      TMA producer : Q once, then K_j / V_j tiles ([128 kv × 128 d] as two 128B-swizzled [128 × 64] boxes)
//   warp 1      MMA issuer   : S = Q·K_jᵀ  (A, B K-major, 8 × UMMA 128×128×16)  → TMEM cols [0,128)
//                              O += P·V_j (A = P read from TMEM, B = V MN-major)  → TMEM cols [128,256)
//   warps 2-5   softmax      : thread = query row (tcgen05.ld 32x32b: lane ↔ TMEM lane): row max, exp2, row sum, P packed
//                              to bf16 and written back over S with tcgen05.st.  The O accumulator is only rescaled when
//                              the running max grew by more than 2^8 (lazy rescale), so the common tile does no TMEM
//                              round trip for O at all.
// Outputs: O (bf16, arbitrary b/s/h strides) and LSE (fp32 [B, H, S], natural log) for the backward.
#include <cuda.h>

#include <cstdlib>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace nxd {

CUtensorMap make_tmap_bf16_strided(const void* ptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                                   uint32_t box_cols, uint32_t box_rows);

namespace fa {

constexpr int BM = 128, BN = 128, HD = 128;
constexpr int kThreads = 192;
constexpr int kPolyPairs = 0;      // of every 8 element pairs, how many use ex2_poly in the forward softmax
constexpr uint32_t kTile = BM * HD * 2;          // 32 KB
constexpr uint32_t kHalf = kTile / 2;            // one [128 × 64] box
constexpr uint32_t kSmemFwd = 6 * kTile + 1024 /*align*/ + 256 /*barriers*/;

// Warp-uniform issue: the whole MMA warp runs the issue loop in convergent control flow and one elected lane executes each
// tcgen05 instruction (helpers tcgen05_mma_f16_e / tcgen05_commit_e in sm100_ptx.cuh explain why; mma_ts is the variant whose
// A operand is read from TMEM — P of the forward, Pᵀ of the backward).
NXD_DEVICE void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
NXD_DEVICE void tcgen05_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
NXD_DEVICE void tcgen05_st_32x16p(uint32_t taddr, const uint32_t* r) {   // first 16 registers of a larger array
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
NXD_DEVICE void tcgen05_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
NXD_DEVICE void tcgen05_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
NXD_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): halves the issue slots of the softmax element math
NXD_DEVICE uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
NXD_DEVICE uint64_t pack2u(uint32_t a, uint32_t b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
NXD_DEVICE void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
NXD_DEVICE uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
NXD_DEVICE uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
NXD_DEVICE uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
NXD_DEVICE uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// TMA coordinates of element (b, head, s, 0): c0 = b*col_b + head*col_h, c1 = b*row_b + head*row_h + s.
struct Coord { int col_b, col_h, row_b, row_h; };

struct FwdParams {
  int S_q, S_kv, H, Hkv, B;
  Coord q, k, v;
  long o_sb, o_ss, o_sh;      // output strides in elements
  float scale, scale_log2;
  int causal;
  float* lse;                 // [B, H, S_q]
};

// exp2 on the FMA pipe (Cody–Waite split + degree-3 minimax on [-0.5, 0.5], rel. error ≈ 1e-4 — far below bf16's 4e-3):
// the MUFU unit does 4 lanes/clk/SMSP, so a 128-wide row costs 1024 MUFU cycles per warp; moving ~30 % of the elements
// here balances MUFU time against issue slots.
NXD_DEVICE float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;                   // 1.5·2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.f);             // f ∈ [-0.5, 0.5]
  float pz = fmaf(f, 0.0550087f, 0.2422104f);      // minimax fit of 2^f, max rel. error 1.0e-4
  pz = fmaf(pz, f, 0.6932829f);
  pz = fmaf(pz, f, 1.0f);
  return __int_as_float(__float_as_int(pz) + (__float_as_int(t) << 23));
}

// Persistent forward: one CTA per SM walks (q-tile, head, batch) work items in round-robin order.
// TMEM: S double buffer [0,256) (P_j is written over S_j in place), O double buffer [256,512).  Tensor-pipe order is
// S_0 S_1 PV_0 S_2 PV_1 S_3 … across item boundaries: the QKᵀ of the next tiles (including the first tiles of the NEXT work
// item, whose Q sits in the second Q buffer) and the PV of the previous tile run while the softmax warps work on the current
// tile, and the epilogue of item i (O_i → HBM) overlaps the first QKᵀ/PV of item i+1 (second O buffer).  The one-time costs
// (TMEM allocation, barrier init, descriptor prefetch, first TMA round trip) are paid once per SM instead of once per tile.
struct FwdItem { int qt, head, b, n_kv; };

// Work order = (batch·head) major, q-tile minor — the order the hardware block scheduler would use, so the CTAs running at any
// moment work on the same few heads and K/V stay in L2 (sorting all items heaviest-first instead made every CTA stream a
// different head and re-read K/V from HBM for every q tile: 4.4 GB instead of 0.27 GB per call, measured 15 % slower).  The q
// tile is rotated by the head index so that a CTA's round-robin slice (stride = grid size) mixes heavy and light causal tiles.
NXD_DEVICE FwdItem fwd_item(int idx, int nq, int n_tiles_kv, const FwdParams& p) {
  FwdItem it;
  const int bh = idx / nq, j = idx % nq;
  it.qt = nq - 1 - (j + bh) % nq;
  it.head = bh % p.H;
  it.b = bh / p.H;
  it.n_kv = p.causal ? min(it.qt + 1, n_tiles_kv) : n_tiles_kv;
  return it;
}
NXD_DEVICE int fwd_item_index(int k) { return k * (int)gridDim.x + (int)blockIdx.x; }

__global__ void __launch_bounds__(kThreads, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
              const __grid_constant__ CUtensorMap tv, __nv_bfloat16* __restrict__ out, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sK = smem_base + 2 * kTile, sV = smem_base + 4 * kTile;   // Q, K, V: two stages each
  const uint32_t bars = smem_base + 6 * kTile;
  // qf[2] qe[2] kf[2] vf[2] ke[2] ve[2] s[2] p[2] pv oe[2] | tmem slot
  const uint32_t bar_qf = bars, bar_qe = bars + 16, bar_kf = bars + 32, bar_vf = bars + 48, bar_ke = bars + 64,
                 bar_ve = bars + 80, bar_s = bars + 96, bar_p = bars + 112, bar_pv = bars + 128, bar_oe = bars + 136,
                 tmem_slot = bars + 152;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.S_q + BM - 1) / BM;
  const int n_tiles_kv = (p.S_kv + BN - 1) / BN;
  const int n_items = nq * p.H * p.B;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qf + 8 * i, 1); mbar_init(bar_qe + 8 * i, 1);
      mbar_init(bar_kf + 8 * i, 1); mbar_init(bar_vf + 8 * i, 1); mbar_init(bar_ke + 8 * i, 1); mbar_init(bar_ve + 8 * i, 1);
      mbar_init(bar_s + 8 * i, 1); mbar_init(bar_p + 8 * i, 128); mbar_init(bar_oe + 8 * i, 128);
    }
    mbar_init(bar_pv, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tcgen05_alloc(tmem_slot, 512);
    tcgen05_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv);
      int g = 0;
      for (int il = 0;; ++il) {
        const int idx = fwd_item_index(il);
        if (idx >= n_items) break;
        const FwdItem it = fwd_item(idx, nq, n_tiles_kv, p);
        const int kvh = it.head / (p.H / p.Hkv);
        const int qb = il & 1;
        mbar_wait(bar_qe + 8 * qb, (uint32_t)(((il >> 1) & 1) ^ 1));
        const int qc = it.b * p.q.col_b + it.head * p.q.col_h, qr = it.b * p.q.row_b + it.head * p.q.row_h + it.qt * BM;
        mbar_expect_tx(bar_qf + 8 * qb, kTile);
        tma_load_2d(sQ + qb * kTile, &tq, bar_qf + 8 * qb, qc, qr);
        tma_load_2d(sQ + qb * kTile + kHalf, &tq, bar_qf + 8 * qb, qc + 64, qr);
        const int kc = it.b * p.k.col_b + kvh * p.k.col_h, kr = it.b * p.k.row_b + kvh * p.k.row_h;
        const int vc = it.b * p.v.col_b + kvh * p.v.col_h, vr = it.b * p.v.row_b + kvh * p.v.row_h;
        for (int j = 0; j < it.n_kv; ++j, ++g) {
          const int st = g & 1;
          const uint32_t ph = (uint32_t)((g >> 1) & 1);
          mbar_wait(bar_ke + 8 * st, ph ^ 1);
          mbar_expect_tx(bar_kf + 8 * st, kTile);
          tma_load_2d(sK + st * kTile, &tk, bar_kf + 8 * st, kc, kr + j * BN);
          tma_load_2d(sK + st * kTile + kHalf, &tk, bar_kf + 8 * st, kc + 64, kr + j * BN);
          mbar_wait(bar_ve + 8 * st, ph ^ 1);
          mbar_expect_tx(bar_vf + 8 * st, kTile);
          tma_load_2d(sV + st * kTile, &tv, bar_vf + 8 * st, vc, vr + j * BN);
          tma_load_2d(sV + st * kTile + kHalf, &tv, bar_vf + 8 * st, vc + 64, vr + j * BN);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // all 32 lanes run this loop; the helpers elect one lane per tcgen05 instruction
      constexpr uint32_t idesc_s = make_idesc(false, false, BM, BN);
      constexpr uint32_t idesc_o = make_idesc(false, true, BM, HD);
      // cursor over the flattened (item, kv-tile) sequence of this CTA
      struct Cur { int il, j, n_kv, g; bool ok; };
      auto start = [&](int il, int g) {
        Cur c{il, 0, 0, g, false};
        const int idx = fwd_item_index(il);
        if (idx < n_items) { c.n_kv = fwd_item(idx, nq, n_tiles_kv, p).n_kv; c.ok = true; }
        return c;
      };
      auto advance = [&](Cur& c) {
        ++c.g;
        if (++c.j == c.n_kv) c = start(c.il + 1, c.g);
      };
      auto issue_s = [&](const Cur& c) {
        const int st = c.g & 1, qb = c.il & 1;
        if (c.j == 0) mbar_wait(bar_qf + 8 * qb, (uint32_t)((c.il >> 1) & 1));
        mbar_wait(bar_kf + 8 * st, (uint32_t)((c.g >> 1) & 1));
        tcgen05_fence_after();
        const uint32_t q_s = sQ + qb * kTile, k_s = sK + st * kTile;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tcgen05_mma_f16_e(tmem_S + st * BN, make_smem_desc(q_s + kb * kHalf + kk * 32, 16, 1024),
                   make_smem_desc(k_s + kb * kHalf + kk * 32, 16, 1024), idesc_s, (kb | kk) ? 1u : 0u);
        tcgen05_commit_e(bar_ke + 8 * st);
        tcgen05_commit_e(bar_s + 8 * st);
        if (c.j == c.n_kv - 1) tcgen05_commit_e(bar_qe + 8 * qb);       // last QKᵀ of the item: its Q buffer is free
      };
      Cur ahead = start(0, 0), cur = ahead;
      if (ahead.ok) { issue_s(ahead); advance(ahead); }
      if (ahead.ok) { issue_s(ahead); advance(ahead); }
      while (cur.ok) {
        const int st = cur.g & 1, ob = cur.il & 1;
        const uint32_t ph = (uint32_t)((cur.g >> 1) & 1);
        mbar_wait(bar_p + 8 * st, ph);
        mbar_wait(bar_vf + 8 * st, ph);
        if (cur.j == 0) mbar_wait(bar_oe + 8 * ob, (uint32_t)(((cur.il >> 1) & 1) ^ 1));   // epilogue of item il-2 left O[ob]
        tcgen05_fence_after();
        const uint32_t v_s = sV + st * kTile;
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          mma_ts(tmem_O + ob * HD, tmem_S + st * BN + k * 8, make_smem_desc(v_s + k * 2048, kHalf, 1024), idesc_o,
                 (cur.j | k) ? 1u : 0u);
        tcgen05_commit_e(bar_ve + 8 * st);
        tcgen05_commit_e(bar_pv);
        if (ahead.ok) { issue_s(ahead); advance(ahead); }        // overwrites P of tile g — behind PV_g in pipe order
        advance(cur);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale_log2;
    int g = 0;
    for (int il = 0;; ++il) {
      const int idx = fwd_item_index(il);
      if (idx >= n_items) break;
      const FwdItem it = fwd_item(idx, nq, n_tiles_kv, p);
      const int qt = it.qt, ob = il & 1;
      const int q_idx = qt * BM + row;
      const uint32_t tO = tmem_O + ob * HD + lane_base;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < it.n_kv; ++j, ++g) {
        const int st = g & 1;
        const uint32_t tS = tmem_S + st * BN + lane_base;
        mbar_wait(bar_s + 8 * st, (uint32_t)((g >> 1) & 1));
        tcgen05_fence_after();
        const int kv0 = j * BN;
        const bool masked = (p.causal && kv0 + BN - 1 > qt * BM) || (kv0 + BN > p.S_kv);
        // ---- the whole S row (128 fp32) comes into registers with four back-to-back TMEM loads and one wait
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tcgen05_ld_32x32(tS + c * 32, sr[c]);
        tcgen05_wait_ld();
        if (masked) {      // diagonal / ragged tile: −inf in place, the exp below turns it into an exact 0
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int kv = kv0 + c * 32 + i;
              if (!(kv < p.S_kv && (!p.causal || kv <= q_idx))) sr[c][i] = 0xff800000u;
            }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 32; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(mx4[c], __uint_as_float(sr[c][i]));
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if (j == 0) {
          m_ref = mx;
        } else {
          const float m_new = fmaxf(m_ref, mx);
          const bool need = (m_new - m_ref) * sl2 > 8.f;
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(bar_pv, (uint32_t)((g - 1) & 1));      // PV of the previous tile retired: O is quiescent
            tcgen05_fence_after();
            const float f = need ? ex2((m_ref - m_new) * sl2) : 1.f;
            if (need) m_ref = m_new;
            l *= f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t r[32];
              tcgen05_ld_32x32(tO + c * 32, r);
              tcgen05_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tcgen05_st_32x32(tO + c * 32, r);
            }
          }
        }
        // ---- P = exp2(s·scale·log2e − m_ref·scale·log2e), packed to bf16 in place and written over the S columns
        const float mb = m_ref * sl2;
        const uint64_t sl2_2 = pack2(sl2, sl2), nmb_2 = pack2(-mb, -mb);
        uint64_t l2acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            unpack2(fma2(pack2u(sr[c][i], sr[c][i + 1]), sl2_2, nmb_2), x0, x1);
            // kPolyPairs of every 8 pairs take the FMA-pipe exp2; 0 = all on MUFU (a lone warp per SMSP is issue-bound, not
            // MUFU-bound, once the polynomial's 8 extra instructions are counted — measured, see profiles/)
            const bool poly = ((i >> 1) & 7) < kPolyPairs;
            const float p0 = poly ? ex2_poly(x0) : ex2(x0);
            const float p1 = poly ? ex2_poly(x1) : ex2(x1);
            l2acc[c] = add2(l2acc[c], pack2(p0, p1));
            sr[c][i >> 1] = pack_bf16(p0, p1);
          }
          tcgen05_st_32x16p(tS + c * 16, sr[c]);
        }
        float l4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { float a, bq; unpack2(l2acc[c], a, bq); l4[c] = a + bq; }
        l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        tcgen05_wait_st();
        tcgen05_fence_before();
        mbar_arrive(bar_p + 8 * st);
      }
      // ---- epilogue: O / l → bf16, LSE (the MMA warp is already working on the next item in the other O buffer)
      mbar_wait(bar_pv, (uint32_t)((g - 1) & 1));
      tcgen05_fence_after();
      const float inv_l = 1.f / l;
      const bool row_ok = q_idx < p.S_q;
      __nv_bfloat16* orow = out + (long)it.b * p.o_sb + (long)q_idx * p.o_ss + (long)it.head * p.o_sh;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tcgen05_ld_32x32(tO + c * 32, r);
        tcgen05_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[v4 * 8 + 0]) * inv_l, __uint_as_float(r[v4 * 8 + 1]) * inv_l);
            o.y = pack_bf16(__uint_as_float(r[v4 * 8 + 2]) * inv_l, __uint_as_float(r[v4 * 8 + 3]) * inv_l);
            o.z = pack_bf16(__uint_as_float(r[v4 * 8 + 4]) * inv_l, __uint_as_float(r[v4 * 8 + 5]) * inv_l);
            o.w = pack_bf16(__uint_as_float(r[v4 * 8 + 6]) * inv_l, __uint_as_float(r[v4 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c * 32 + v4 * 8) = o;
          }
        }
      }
      if (row_ok) p.lse[((long)it.b * p.H + it.head) * p.S_q + q_idx] = m_ref * p.scale + __logf(l);
      tcgen05_fence_before();
      mbar_arrive(bar_oe + 8 * ob);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tcgen05_dealloc(tmem_base, 512);
}

// ====================================================================================================================
// Backward.  One CTA per (128-row KV tile, kv head, batch); it keeps K/V resident in smem and dK/dV resident in TMEM and
// walks the 64-row query tiles (× the GQA group) that can see this KV tile.  Everything is computed transposed so that the
// softmax threads own a *kv* row and P/dS land directly in the operand layouts the next GEMMs need:
//   Sᵀ  = K·Qᵀ        (A = K  K-major, B = Q  K-major)            → TMEM [128 kv × 64 q]
//   dPᵀ = V·dOᵀ       (A = V  K-major, B = dO K-major)            → TMEM [128 kv × 64 q]
//   Pᵀ = exp2(Sᵀ·c − lse₂[q]) ;  dSᵀ = Pᵀ ∘ (dPᵀ·scale − δ[q]·scale)   (thread = kv row; lse₂/δ broadcast from smem)
//   dV += Pᵀ·dO       (A = Pᵀ from TMEM (in place over Sᵀ), B = dO MN-major)
//   dK += dSᵀ·Q       (A = dSᵀ smem K-major,               B = Q  MN-major)
//   dQᵀ = Kᵀ·dSᵀ      (A = K MN-major, B = dSᵀ MN-major)          → TMEM [128 d × 64 q], drained by a second warpgroup
//                      into smem and added to the fp32 dQ accumulator with one bulk async reduce (no per-element atomics).
// warp 0 TMA, warp 1 MMA issue, warps 2-5 softmax/dS (+ dK/dV epilogue), warps 6-9 dQ drain.
constexpr int BQ = 64;
constexpr int kThreadsBwd = 320;
constexpr int kQStages = 3;
constexpr uint32_t kQTile = BQ * HD * 2;                        // 16 KB
constexpr uint32_t kOffK = 0, kOffV = kTile, kOffQ = 2 * kTile, kOffdO = kOffQ + kQStages * kQTile,
                   kOffdS = kOffdO + kQStages * kQTile, kOffdQ = kOffdS + 2 * kQTile, kOffStats = kOffdQ + 16384,
                   kOffBars = kOffStats + kQStages * 512;
constexpr uint32_t kSmemBwd = kOffBars + 256 + 1024;

struct BwdParams {
  int S_q, S_kv, H, Hkv, B, S_pad;
  Coord q, k, v, g;
  long dk_sb, dk_ss, dk_sh, dv_sb, dv_ss, dv_sh;
  float scale, scale_log2;
  int causal;
  const float* stats;          // [2][B*H*S_pad (+pad)] : −lse·log2e , −δ·scale
  long stats_stride;
  float* dq_acc;               // [B, H, S_pad, 128] fp32
  int debug;                   // NXD_FA_DEBUG bits (perf triage only): 1 no bulk reduce, 2 no dQ staging, 4 no softmax math
  unsigned* turnstile;         // DET only: [B, H, ceil(S_q / BQ)] contributions already added to each dQ tile
};

NXD_DEVICE void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
NXD_DEVICE void bulk_reduce_add_f32(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
}
NXD_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
NXD_DEVICE void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
NXD_DEVICE void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
NXD_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
NXD_DEVICE void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Softmax-side math of one backward tile, split in two phases so the exp work (needs only Sᵀ, which the tensor pipe delivered a
// full iteration ago) runs while dPᵀ of the same tile is still being computed:
//   phase 1: Pᵀ = exp2(Sᵀ·c − lse₂[q])  → fp32 in registers (pf) + bf16 pairs packed in place into sv (element pair (i,i+1) → i/2)
//   phase 2: dSᵀ = Pᵀ ∘ (dPᵀ·scale − δ[q]·scale) → bf16 pairs packed in place into dpv
// The stats rows hold −lse·log2e and −δ·scale so both affine steps are single packed FMAs; MASKED is a template parameter so
// the common tile has no per-element branches or selects.
template <bool MASKED>
NXD_DEVICE void bwd_p_math(uint32_t (&sv)[2][32], float (&pf)[2][32], uint32_t stats_s, float sl2, int kv, int q0, int S_q,
                           int S_kv, int causal) {
  const uint64_t sl2_2 = pack2(sl2, sl2);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 l2;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(l2.x), "=f"(l2.y), "=f"(l2.z), "=f"(l2.w)
                   : "r"(stats_s + (h * 32 + i) * 4));
      const uint64_t nl2[2] = {pack2(l2.x, l2.y), pack2(l2.z, l2.w)};
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        float x0, x1;
        unpack2(fma2(pack2u(sv[h][i + 2 * e2], sv[h][i + 2 * e2 + 1]), sl2_2, nl2[e2]), x0, x1);
        float p0 = ex2(x0), p1 = ex2(x1);
        if constexpr (MASKED) {
          const int q = q0 + h * 32 + i + 2 * e2;
          p0 = (kv < S_kv && q < S_q && (!causal || kv <= q)) ? p0 : 0.f;          // selects; also kills NaN padding
          p1 = (kv < S_kv && q + 1 < S_q && (!causal || kv <= q + 1)) ? p1 : 0.f;
        }
        pf[h][i + 2 * e2] = p0;
        pf[h][i + 2 * e2 + 1] = p1;
      }
      sv[h][i >> 1] = pack_bf16(pf[h][i], pf[h][i + 1]);
      sv[h][(i >> 1) + 1] = pack_bf16(pf[h][i + 2], pf[h][i + 3]);
    }
  }
}

template <bool MASKED>
NXD_DEVICE void bwd_ds_math(uint32_t (&dpv)[2][32], const float (&pf)[2][32], uint32_t stats_s, float sc) {
  const uint64_t sc_2 = pack2(sc, sc);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float4 dl;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(dl.x), "=f"(dl.y), "=f"(dl.z), "=f"(dl.w)
                   : "r"(stats_s + 256 + (h * 32 + i) * 4));
      const uint64_t ndl[2] = {pack2(dl.x, dl.y), pack2(dl.z, dl.w)};
      float dsv[4];
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        float d0, d1;
        unpack2(mul2(pack2(pf[h][i + 2 * e2], pf[h][i + 2 * e2 + 1]),
                     fma2(pack2u(dpv[h][i + 2 * e2], dpv[h][i + 2 * e2 + 1]), sc_2, ndl[e2])), d0, d1);
        if constexpr (MASKED) {          // P is already exactly 0 where masked, but 0·NaN (padding garbage) must not leak
          d0 = pf[h][i + 2 * e2] == 0.f ? 0.f : d0;
          d1 = pf[h][i + 2 * e2 + 1] == 0.f ? 0.f : d1;
        }
        dsv[2 * e2] = d0; dsv[2 * e2 + 1] = d1;
      }
      dpv[h][i >> 1] = pack_bf16(dsv[0], dsv[1]);
      dpv[h][(i >> 1) + 1] = pack_bf16(dsv[2], dsv[3]);
    }
  }
}

// DET = deterministic dQ: the contributions of the K/V tiles to one dQ tile are added in K/V-tile order through a turnstile
// counter per (batch, head, query tile) — CTA n of a (batch, KV head) column waits until n earlier tiles have landed, adds its
// own with the same bulk reduce, waits for that reduce to COMPLETE and releases the next.  CTAs are dispatched in increasing
// blockIdx.x, so the CTA being waited for is always resident or done.  Slower (the drain of a tile is serialised across the
// column), bit-reproducible.  DET = false compiles to exactly the kernel without the option.
template <bool DET>
__global__ void __launch_bounds__(kThreadsBwd, 1)
fa_bwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
              const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap tg,
              __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = base + kOffK, sV = base + kOffV, sQ = base + kOffQ, sdO = base + kOffdO, sdS = base + kOffdS,
                 sdQ = base + kOffdQ, sStats = base + kOffStats, bars = base + kOffBars;
  // barriers: kv | qf[3] | qe[3] | s[2] | dp | p | dq | dqr | acc | tmem slot
  const uint32_t bar_kv = bars, bar_qf = bars + 8, bar_qe = bars + 32, bar_s = bars + 56, bar_dp = bars + 72,
                 bar_p = bars + 80, bar_dq = bars + 88, bar_dqr = bars + 96, bar_acc = bars + 104, tmem_slot = bars + 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int group = p.H / p.Hkv;
  const int nq = (p.S_q + BQ - 1) / BQ;
  const int t0 = p.causal ? (n * BN) / BQ : 0;
  const int per_head = nq - t0;
  const int n_iter = per_head * group;
  const int kv0 = n * BN;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv, 1);
    for (int i = 0; i < kQStages; ++i) { mbar_init(bar_qf + 8 * i, 1); mbar_init(bar_qe + 8 * i, 1); }
    mbar_init(bar_s, 1); mbar_init(bar_s + 8, 1);
    mbar_init(bar_dp, 1); mbar_init(bar_p, 128); mbar_init(bar_dq, 1); mbar_init(bar_dqr, 128);
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tcgen05_alloc(tmem_slot, 512);
    tcgen05_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // TMEM columns: dK [0,128) | dV [128,256) | Sᵀ/Pᵀ double buffer [256,384) | dPᵀ [384,448) | dQᵀ [448,512)
  const uint32_t tmem_dK = tmem_base, tmem_dV = tmem_base + 128, tmem_ST = tmem_base + 256, tmem_dP = tmem_base + 384,
                 tmem_dQ = tmem_base + 448;

  if (warp == 0) {
    if (lane == 0 && n_iter > 0) {
      prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv); prefetch_tmap(&tg);
      const int kc = b * p.k.col_b + kvh * p.k.col_h, kr = b * p.k.row_b + kvh * p.k.row_h + kv0;
      const int vc = b * p.v.col_b + kvh * p.v.col_h, vr = b * p.v.row_b + kvh * p.v.row_h + kv0;
      mbar_expect_tx(bar_kv, 2 * kTile);
      tma_load_2d(sK, &tk, bar_kv, kc, kr);
      tma_load_2d(sK + kHalf, &tk, bar_kv, kc + 64, kr);
      tma_load_2d(sV, &tv, bar_kv, vc, vr);
      tma_load_2d(sV + kHalf, &tv, bar_kv, vc + 64, vr);
      for (int it = 0; it < n_iter; ++it) {
        const int qs = it % kQStages;
        const uint32_t ph = (uint32_t)((it / kQStages) & 1);
        const int head = kvh * group + it / per_head, t = t0 + it % per_head;
        mbar_wait(bar_qe + 8 * qs, ph ^ 1);
        const uint32_t full = bar_qf + 8 * qs;
        mbar_expect_tx(full, 2 * kQTile + 512);
        const int qc = b * p.q.col_b + head * p.q.col_h, qr = b * p.q.row_b + head * p.q.row_h + t * BQ;
        const int gc = b * p.g.col_b + head * p.g.col_h, gr = b * p.g.row_b + head * p.g.row_h + t * BQ;
        tma_load_2d(sQ + qs * kQTile, &tq, full, qc, qr);
        tma_load_2d(sQ + qs * kQTile + kQTile / 2, &tq, full, qc + 64, qr);
        tma_load_2d(sdO + qs * kQTile, &tg, full, gc, gr);
        tma_load_2d(sdO + qs * kQTile + kQTile / 2, &tg, full, gc + 64, gr);
        const float* st = p.stats + ((long)b * p.H + head) * p.S_pad + t * BQ;
        bulk_load_1d(sStats + qs * 512, st, 256, full);
        bulk_load_1d(sStats + qs * 512 + 256, st + p.stats_stride, 256, full);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (n_iter > 0) {   // all 32 lanes run the issue loop; one elected lane executes each tcgen05 instruction
      constexpr uint32_t idesc_s = make_idesc(false, false, 128, BQ);      // Sᵀ, dPᵀ
      constexpr uint32_t idesc_acc = make_idesc(false, true, 128, HD);     // dK, dV
      constexpr uint32_t idesc_dq = make_idesc(true, true, 128, BQ);       // dQᵀ
      mbar_wait(bar_kv, 0);
      // Issue order per iteration `it` (everything on the tensor pipe retires in this order):
      //   … Sᵀ_{it+1} already issued …  | wait softmax(it) | dPᵀ_{it+1} | dK_it dV_it | dQᵀ_it | Sᵀ_{it+2}
      // so while the softmax warps work on tile it+1 the pipe runs dK/dV/dQ of tile it and Sᵀ of tile it+2.
      auto issue_st = [&](int it) {
        const int qs = it % kQStages, st = it & 1;
        mbar_wait(bar_qf + 8 * qs, (uint32_t)((it / kQStages) & 1));
        tcgen05_fence_after();
        const uint32_t q_s = sQ + qs * kQTile;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tcgen05_mma_f16_e(tmem_ST + st * BQ, make_smem_desc(sK + kb * kHalf + kk * 32, 16, 1024),
                            make_smem_desc(q_s + kb * (kQTile / 2) + kk * 32, 16, 1024), idesc_s, (kb | kk) ? 1u : 0u);
        tcgen05_commit_e(bar_s + 8 * st);
      };
      auto issue_dp = [&](int it) {
        const int qs = it % kQStages;
        mbar_wait(bar_qf + 8 * qs, (uint32_t)((it / kQStages) & 1));
        tcgen05_fence_after();
        const uint32_t g_s = sdO + qs * kQTile;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tcgen05_mma_f16_e(tmem_dP, make_smem_desc(sV + kb * kHalf + kk * 32, 16, 1024),
                            make_smem_desc(g_s + kb * (kQTile / 2) + kk * 32, 16, 1024), idesc_s, (kb | kk) ? 1u : 0u);
        tcgen05_commit_e(bar_dp);
      };
      issue_st(0);
      issue_dp(0);
      if (n_iter > 1) issue_st(1);
      for (int it = 0; it < n_iter; ++it) {
        const int qs = it % kQStages, st = it & 1;
        mbar_wait(bar_p, (uint32_t)(it & 1));             // Pᵀ_it in TMEM, dSᵀ_it in smem, dPᵀ region consumed
        tcgen05_fence_after();
        if (it + 1 < n_iter) issue_dp(it + 1);
        const uint32_t q_s = sQ + qs * kQTile, g_s = sdO + qs * kQTile, ds_s = sdS + st * kQTile;
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk)
          tcgen05_mma_f16_e(tmem_dK, make_smem_desc(ds_s + kk * 32, 16, 1024), make_smem_desc(q_s + kk * 2048, kQTile / 2, 1024),
                          idesc_acc, (it | kk) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk)
          mma_ts(tmem_dV, tmem_ST + st * BQ + kk * 8, make_smem_desc(g_s + kk * 2048, kQTile / 2, 1024), idesc_acc,
                         (it | kk) ? 1u : 0u);
        tcgen05_commit_e(bar_qe + 8 * qs);
        if (it >= 1) {
          mbar_wait(bar_dqr, (uint32_t)((it - 1) & 1));   // dQᵀ_{it-1} has been read out of TMEM
          tcgen05_fence_after();
        }
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          tcgen05_mma_f16_e(tmem_dQ, make_smem_desc(sK + kk * 2048, kHalf, 1024), make_smem_desc(ds_s + kk * 2048, kHalf, 1024),
                          idesc_dq, kk ? 1u : 0u);
        tcgen05_commit_e(bar_dq);
        if (it + 2 < n_iter) issue_st(it + 2);            // overwrites Pᵀ_it — after dV_it in pipe order
      }
      tcgen05_commit_e(bar_acc);
    }
    __syncwarp();
  } else if (warp < 6) {
    // ===== softmax / dS warps: thread = kv row =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int kv = kv0 + row;
    const float sl2 = p.scale_log2, sc = p.scale;
    for (int it = 0; it < n_iter; ++it) {
      const int qs = it % kQStages, st = it & 1;
      const int t = t0 + it % per_head;
      const int q0 = t * BQ;
      mbar_wait(bar_qf + 8 * qs, (uint32_t)((it / kQStages) & 1));     // stats (TMA-written) visible to this thread
      mbar_wait(bar_s + 8 * st, (uint32_t)((it >> 1) & 1));
      tcgen05_fence_after();
      const bool masked = (p.causal && q0 < kv0 + BN - 1) || (q0 + BQ > p.S_q) || (kv0 + BN > p.S_kv);
      const uint32_t stats_s = sStats + qs * 512;
      const uint32_t ds_row = sdS + st * kQTile + row * 128;
      // ---- phase 1 (needs only Sᵀ): Pᵀ, written back over Sᵀ in TMEM
      uint32_t sv[2][32];
      float pf[2][32];
      tcgen05_ld_32x32(tmem_ST + st * BQ + lane_base, sv[0]);
      tcgen05_ld_32x32(tmem_ST + st * BQ + lane_base + 32, sv[1]);
      tcgen05_wait_ld();
      if (masked) bwd_p_math<true>(sv, pf, stats_s, sl2, kv, q0, p.S_q, p.S_kv, p.causal);
      else bwd_p_math<false>(sv, pf, stats_s, sl2, kv, q0, p.S_q, p.S_kv, p.causal);
      tcgen05_st_32x16p(tmem_ST + st * BQ + lane_base, sv[0]);
      tcgen05_st_32x16p(tmem_ST + st * BQ + lane_base + 16, sv[1]);
      // ---- phase 2 (dPᵀ of this tile was issued while phase 1 ran): dSᵀ → smem
      mbar_wait(bar_dp, (uint32_t)(it & 1));
      tcgen05_fence_after();
      uint32_t dpv[2][32];
      tcgen05_ld_32x32(tmem_dP + lane_base, dpv[0]);
      tcgen05_ld_32x32(tmem_dP + lane_base + 32, dpv[1]);
      tcgen05_wait_ld();
      if (!(p.debug & 4)) {
        if (masked) bwd_ds_math<true>(dpv, pf, stats_s, sc);
        else bwd_ds_math<false>(dpv, pf, stats_s, sc);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // dSᵀ row → smem, 128B-swizzled: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t chunk = (uint32_t)(h * 4 + c) ^ (uint32_t)(row & 7);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ds_row + chunk * 16), "r"(dpv[h][c * 4]),
                       "r"(dpv[h][c * 4 + 1]), "r"(dpv[h][c * 4 + 2]), "r"(dpv[h][c * 4 + 3]) : "memory");
        }
      }
      tcgen05_wait_st();
      fence_async_smem();
      tcgen05_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- epilogue: dK, dV
    if (n_iter > 0) {
      mbar_wait(bar_acc, 0);
      tcgen05_fence_after();
    }
    const bool row_ok = kv < p.S_kv;
    __nv_bfloat16* dk_row = dk + (long)b * p.dk_sb + (long)kv * p.dk_ss + (long)kvh * p.dk_sh;
    __nv_bfloat16* dv_row = dv + (long)b * p.dv_sb + (long)kv * p.dv_ss + (long)kvh * p.dv_sh;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* orow = which ? dv_row : dk_row;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        if (n_iter > 0) {
          tcgen05_ld_32x32((which ? tmem_dV : tmem_dK) + lane_base + c * 32, r);
          tcgen05_wait_ld();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0;
        }
        if (row_ok) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r[v4 * 8 + 0]), __uint_as_float(r[v4 * 8 + 1]));
            o.y = pack_bf16(__uint_as_float(r[v4 * 8 + 2]), __uint_as_float(r[v4 * 8 + 3]));
            o.z = pack_bf16(__uint_as_float(r[v4 * 8 + 4]), __uint_as_float(r[v4 * 8 + 5]));
            o.w = pack_bf16(__uint_as_float(r[v4 * 8 + 6]), __uint_as_float(r[v4 * 8 + 7]));
            *reinterpret_cast<uint4*>(orow + c * 32 + v4 * 8) = o;
          }
        }
      }
    }
  } else {
    // ===== dQ drain warps: thread = head-dim index d (TMEM lane), columns = q =====
    const int quarter = warp & 3;
    const int d = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int tid = threadIdx.x - 192;
    for (int it = 0; it < n_iter; ++it) {
      const int st = it & 1;
      const int head = kvh * group + it / per_head, t = t0 + it % per_head;
      mbar_wait(bar_dq, (uint32_t)(it & 1));
      tcgen05_fence_after();
      float* gdst = p.dq_acc + (((long)b * p.H + head) * p.S_pad + (long)t * BQ) * HD;
      uint32_t r[2][32];
      tcgen05_ld_32x32(tmem_dQ + lane_base, r[0]);
      tcgen05_ld_32x32(tmem_dQ + lane_base + 32, r[1]);
      tcgen05_wait_ld();
      tcgen05_fence_before();
      mbar_arrive(bar_dqr);                     // the MMA warp may overwrite dQᵀ with the next tile right away
      if (p.debug & 2) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (tid == 0) bulk_wait_read0();        // previous bulk reduce has finished reading the staging tile
        named_bar(1, 128);
#pragma unroll
        for (int c = 0; c < 32; ++c)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sdQ + c * 512 + d * 4), "r"(r[h][c]) : "memory");
        fence_async_smem();
        named_bar(1, 128);
        if (tid == 0 && !(p.debug & 1)) {
          if constexpr (DET) {
            if (h == 0) {                           // my turn: the n earlier K/V tiles of this column have been added
              const unsigned* turn = p.turnstile + ((long)b * p.H + head) * nq + t;
              unsigned seen, spins = 0;
              unsigned long long t_start = 0;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(turn) : "memory");
                if (seen < (unsigned)n && (++spins & 0x3ffu) == 0) {
                  unsigned long long now;
                  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                  if (t_start == 0) t_start = now;
                  else if (now - t_start > 10000000000ull) __trap();          // an earlier tile never arrived
                }
              } while (seen < (unsigned)n);
            }
          }
          bulk_reduce_add_f32(gdst + (long)h * 32 * HD, sdQ, 16384);
          bulk_commit();
          if constexpr (DET) {
            if (h == 1) {                           // both halves of the tile are in memory before the next CTA may add
              bulk_wait0();
              __threadfence();
              unsigned* turn = p.turnstile + ((long)b * p.H + head) * nq + t;
              asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(turn), "r"(1u) : "memory");
            }
          }
        }
      }
    }
    if (tid == 0) bulk_wait0();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tcgen05_dealloc(tmem_base, 512);
}

// δ·scale and lse·log2e per (b, h, s): one warp per row of 128 elements.
__global__ void fa_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ go,
                                   const float* __restrict__ lse, const float* __restrict__ dlse, float* __restrict__ stats,
                                   long stats_stride, int B, int S,
                                   int H, int S_pad, long o_sb, long o_ss, long o_sh, long g_sb, long g_ss, long g_sh,
                                   float scale) {
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long)B * H * S) return;
  const int s = (int)(w % S), h = (int)((w / S) % H), b = (int)(w / ((long)S * H));
  const uint2 ov = *reinterpret_cast<const uint2*>(o + b * o_sb + s * o_ss + h * o_sh + lane * 4);
  const uint2 gv = *reinterpret_cast<const uint2*>(go + b * g_sb + s * g_ss + h * g_sh + lane * 4);
  const __nv_bfloat16* oe = reinterpret_cast<const __nv_bfloat16*>(&ov);
  const __nv_bfloat16* ge = reinterpret_cast<const __nv_bfloat16*>(&gv);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) acc += __bfloat162float(oe[i]) * __bfloat162float(ge[i]);
  acc = warp_sum(acc);
  if (lane == 0) {
    const long idx = ((long)b * H + h) * S_pad + s;
    stats[idx] = -lse[((long)b * H + h) * S + s] * 1.4426950408889634f;
    // a gradient flowing into LSE itself (ring attention merges blocks through their LSEs) adds g_lse·P to dS, i.e. it
    // simply shifts δ: dS = P∘(dP − (δ − g_lse))
    const float g_lse = dlse ? dlse[((long)b * H + h) * S + s] : 0.f;
    stats[stats_stride + idx] = -(acc - g_lse) * scale;
  }
}

// fp32 dQ accumulator [B,H,S_pad,128] → bf16 dq with arbitrary (b, s, h) strides; 8 elements per thread.
__global__ void fa_bwd_dq_out_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int B, int S, int H,
                                     int S_pad, long sb, long ss, long sh) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;        // over B*H*S*16
  if (i >= (long)B * H * S * 16) return;
  const int c = (int)(i & 15);
  const long r = i >> 4;
  const int s = (int)(r % S), h = (int)((r / S) % H), b = (int)(r / ((long)S * H));
  const float4* src = reinterpret_cast<const float4*>(acc + (((long)b * H + h) * S_pad + s) * HD + c * 8);
  const float4 x = src[0], y = src[1];
  uint4 o;
  o.x = pack_bf16(x.x, x.y); o.y = pack_bf16(x.z, x.w); o.z = pack_bf16(y.x, y.y); o.w = pack_bf16(y.z, y.w);
  *reinterpret_cast<uint4*>(dq + b * sb + s * ss + h * sh + c * 8) = o;
}

// ---- host side -----------------------------------------------------------------------------------------------------
// A [B, S, H, 128] bf16 view with strides (sb, ss, sh, 1) becomes a 2-D TMA tensor whose rows are `ss` apart; batch and
// head offsets fold into the row coordinate when they are multiples of ss, else into the column coordinate.
struct View { const void* ptr; int B, S, H; long sb, ss, sh; };

static CUtensorMap view_tmap(const View& v, Coord& c, uint32_t box_rows = 128) {
  long rows = v.S, cols = HD;
  c = Coord{0, 0, 0, 0};
  auto fold = [&](long stride, int n, int& col_mul, int& row_mul) {
    if (n == 1) return;
    if (stride >= v.ss && stride % v.ss == 0) { row_mul = (int)(stride / v.ss); rows += (long)(n - 1) * row_mul; }
    else { col_mul = (int)stride; cols += (long)(n - 1) * stride; }
  };
  fold(v.sb, v.B, c.col_b, c.row_b);
  fold(v.sh, v.H, c.col_h, c.row_h);
  if (cols > v.ss && rows > 1) nxd_throw("attention: unsupported q/k/v strides", __FILE__, __LINE__);
  if ((v.ss * 2) % 16 || ((uintptr_t)v.ptr % 16)) nxd_throw("attention: q/k/v must be 16-byte aligned", __FILE__, __LINE__);
  return make_tmap_bf16_strided(v.ptr, (uint64_t)rows, (uint64_t)cols, (uint64_t)v.ss, 64, box_rows);
}

}  // namespace fa

void flash_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int S_q, int S_kv, int H,
                    int Hkv, const long* qs, const long* ks, const long* vs, const long* os, float scale, bool causal,
                    cudaStream_t st) {
  using namespace fa;
  if (causal && S_q != S_kv) nxd_throw("attention: causal needs S_q == S_kv", __FILE__, __LINE__);
  FwdParams p;
  p.S_q = S_q; p.S_kv = S_kv; p.H = H; p.Hkv = Hkv; p.B = B;
  const CUtensorMap tq = view_tmap(View{q, B, S_q, H, qs[0], qs[1], qs[2]}, p.q);
  const CUtensorMap tk = view_tmap(View{k, B, S_kv, Hkv, ks[0], ks[1], ks[2]}, p.k);
  const CUtensorMap tv = view_tmap(View{v, B, S_kv, Hkv, vs[0], vs[1], vs[2]}, p.v);
  p.o_sb = os[0]; p.o_ss = os[1]; p.o_sh = os[2];
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal ? 1 : 0;
  p.lse = lse;
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(fa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd));
    configured = true;
  }
  const int n_items = ((S_q + BM - 1) / BM) * H * B;
  static int sms = 0;
  if (!sms) { int dev; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  fa_fwd_kernel<<<n_items < sms ? n_items : sms, kThreads, kSmemFwd, st>>>(tq, tk, tv, (__nv_bfloat16*)out, p);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void flash_attn_bwd(const void* go, const void* q, const void* k, const void* v, const void* o, const float* lse,
                    const float* dlse, void* dq,
                    void* dk, void* dv, float* stats, float* dq_acc, int B, int S_q, int S_kv, int H, int Hkv, int S_pad,
                    const long* gs, const long* qs, const long* ks, const long* vs, const long* os, const long* dqs,
                    const long* dks, const long* dvs, float scale, bool causal, cudaStream_t st) {
  using namespace fa;
  if (causal && S_q != S_kv) nxd_throw("attention: causal needs S_q == S_kv", __FILE__, __LINE__);
  BwdParams p;
  p.S_q = S_q; p.S_kv = S_kv; p.H = H; p.Hkv = Hkv; p.B = B; p.S_pad = S_pad;
  const CUtensorMap tq = view_tmap(View{q, B, S_q, H, qs[0], qs[1], qs[2]}, p.q, BQ);
  const CUtensorMap tg = view_tmap(View{go, B, S_q, H, gs[0], gs[1], gs[2]}, p.g, BQ);
  const CUtensorMap tk = view_tmap(View{k, B, S_kv, Hkv, ks[0], ks[1], ks[2]}, p.k);
  const CUtensorMap tv = view_tmap(View{v, B, S_kv, Hkv, vs[0], vs[1], vs[2]}, p.v);
  p.dk_sb = dks[0]; p.dk_ss = dks[1]; p.dk_sh = dks[2];
  p.dv_sb = dvs[0]; p.dv_ss = dvs[1]; p.dv_sh = dvs[2];
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal ? 1 : 0;
  p.stats = stats; p.stats_stride = (long)B * H * S_pad + 64;
  p.dq_acc = dq_acc;
  { const char* e = getenv("NXD_FA_DEBUG"); p.debug = e ? atoi(e) : 0; }
  // NXD_FA_DETERMINISTIC=1: ordered dQ accumulation (see fa_bwd_kernel<true>); the turnstile counters live in a per-device
  // buffer that only grows (debug mode: allocation is synchronous, not meant for graph capture)
  bool det = false;
  { const char* e = getenv("NXD_FA_DETERMINISTIC"); det = e && atoi(e) != 0; }
  p.turnstile = nullptr;
  if (det) {
    static unsigned* turn[64] = {nullptr};
    static size_t cap[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t need = (size_t)B * H * ((S_q + BQ - 1) / BQ) * sizeof(unsigned);
    if (cap[dev & 63] < need) {
      if (turn[dev & 63]) { NXD_CUDA_CHECK(cudaDeviceSynchronize()); cudaFree(turn[dev & 63]); }
      NXD_CUDA_CHECK(cudaMalloc(&turn[dev & 63], need));
      cap[dev & 63] = need;
    }
    p.turnstile = turn[dev & 63];
    NXD_CUDA_CHECK(cudaMemsetAsync(p.turnstile, 0, need, st));
  }
  NXD_CUDA_CHECK(cudaMemsetAsync(dq_acc, 0, (size_t)B * H * S_pad * HD * sizeof(float), st));
  {
    const long warps = (long)B * H * S_q;
    const int threads = 256;
    const long blocks = (warps * 32 + threads - 1) / threads;
    fa_bwd_prep_kernel<<<(unsigned)blocks, threads, 0, st>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)go, lse, dlse, stats,
                                                            p.stats_stride, B, S_q, H, S_pad, os[0], os[1], os[2], gs[0],
                                                            gs[1], gs[2], scale);
  }
  static bool configured = false;
  if (!configured) {
    NXD_CUDA_CHECK(cudaFuncSetAttribute(fa_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd));
    NXD_CUDA_CHECK(cudaFuncSetAttribute(fa_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd));
    configured = true;
  }
  dim3 grid((S_kv + BN - 1) / BN, Hkv, B);
  if (det) fa_bwd_kernel<true><<<grid, kThreadsBwd, kSmemBwd, st>>>(tq, tk, tv, tg, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, p);
  else fa_bwd_kernel<false><<<grid, kThreadsBwd, kSmemBwd, st>>>(tq, tk, tv, tg, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, p);
  {
    const long n = (long)B * H * S_q * 16;
    fa_bwd_dq_out_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dq, B, S_q, H, S_pad, dqs[0],
                                                                     dqs[1], dqs[2]);
  }
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
