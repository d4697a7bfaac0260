// Flash attention for sm_100a (head_dim 128, bf16): tcgen05 QKᵀ / PV with TMEM accumulators, TMA-staged operands,
// online softmax with one thread per query row.  Role parity: reference kernels/flash_attn.py:162-212
// (nki_flash_attn_func → flash_fwd / flash_attn_bwd) and modules/attention call sites.
//
// Forward, one CTA per (128-query tile, head, batch); two CTAs are resident per SM so the softmax of one overlaps the
// tensor-core work of the other (TMEM: 2 × 256 columns):
//   warp 0      TMA producer : Q once, then K_j / V_j tiles ([128 kv × 128 d] as two 128B-swizzled [128 × 64] boxes)
//   warp 1      MMA issuer   : S = Q·K_jᵀ  (A, B K-major, 8 × UMMA 128×128×16)  → TMEM cols [0,128)
//                              O += P·V_j (A = P read from TMEM, B = V MN-major)  → TMEM cols [128,256)
//   warps 2-5   softmax      : thread = query row (tcgen05.ld 32x32b: lane ↔ TMEM lane): row max, exp2, row sum, P packed
//                              to bf16 and written back over S with tcgen05.st.  The O accumulator is only rescaled when
//                              the running max grew by more than 2^8 (lazy rescale), so the common tile does no TMEM
//                              round trip for O at all.
// Outputs: O (bf16, arbitrary b/s/h strides) and LSE (fp32 [B, H, S], natural log) for the backward.
#include <cuda.h>

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
constexpr uint32_t kTile = BM * HD * 2;          // 32 KB
constexpr uint32_t kHalf = kTile / 2;            // one [128 × 64] box
constexpr uint32_t kSmemFwd = 3 * kTile + 1024 /*align*/ + 128 /*barriers*/;

NXD_DEVICE void tcgen05_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
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

__global__ void __launch_bounds__(kThreads, 2)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
              const __grid_constant__ CUtensorMap tv, __nv_bfloat16* __restrict__ out, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base, sK = smem_base + kTile, sV = smem_base + 2 * kTile;
  const uint32_t bars = smem_base + 3 * kTile;
  const uint32_t bar_q = bars, bar_kf = bars + 8, bar_vf = bars + 16, bar_ke = bars + 24, bar_ve = bars + 32,
                 bar_s = bars + 40, bar_p = bars + 48, bar_o = bars + 56, tmem_slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;    // heavy (late) query tiles first
  const int head = blockIdx.y, b = blockIdx.z;
  const int kvh = head / (p.H / p.Hkv);
  const int n_tiles_kv = (p.S_kv + BN - 1) / BN;
  const int n_kv = p.causal ? min(qt + 1, n_tiles_kv) : n_tiles_kv;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_kf, 1); mbar_init(bar_vf, 1); mbar_init(bar_ke, 1); mbar_init(bar_ve, 1);
    mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tcgen05_alloc(tmem_slot, 256);
    tcgen05_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tq); prefetch_tmap(&tk); prefetch_tmap(&tv);
      const int qc = b * p.q.col_b + head * p.q.col_h, qr = b * p.q.row_b + head * p.q.row_h + qt * BM;
      mbar_expect_tx(bar_q, kTile);
      tma_load_2d(sQ, &tq, bar_q, qc, qr);
      tma_load_2d(sQ + kHalf, &tq, bar_q, qc + 64, qr);
      const int kc = b * p.k.col_b + kvh * p.k.col_h, kr = b * p.k.row_b + kvh * p.k.row_h;
      const int vc = b * p.v.col_b + kvh * p.v.col_h, vr = b * p.v.row_b + kvh * p.v.row_h;
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = (uint32_t)(j & 1);
        mbar_wait(bar_ke, ph ^ 1);
        mbar_expect_tx(bar_kf, kTile);
        tma_load_2d(sK, &tk, bar_kf, kc, kr + j * BN);
        tma_load_2d(sK + kHalf, &tk, bar_kf, kc + 64, kr + j * BN);
        mbar_wait(bar_ve, ph ^ 1);
        mbar_expect_tx(bar_vf, kTile);
        tma_load_2d(sV, &tv, bar_vf, vc, vr + j * BN);
        tma_load_2d(sV + kHalf, &tv, bar_vf, vc + 64, vr + j * BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(false, false, BM, BN);
      constexpr uint32_t idesc_o = make_idesc(false, true, BM, HD);
      mbar_wait(bar_q, 0);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = (uint32_t)(j & 1);
        mbar_wait(bar_kf, ph);
        tcgen05_fence_after();
        // S_j overwrites P_{j-1}; tcgen05.mma ops retire in issue order, so PV_{j-1} has consumed P before this lands
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            tcgen05_mma_f16(tmem_S, make_smem_desc(sQ + kb * kHalf + kk * 32, 16, 1024),
                            make_smem_desc(sK + kb * kHalf + kk * 32, 16, 1024), idesc_s, (kb | kk) ? 1u : 0u);
        tcgen05_commit(bar_ke);
        tcgen05_commit(bar_s);
        mbar_wait(bar_p, ph);
        mbar_wait(bar_vf, ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          tcgen05_mma_ts(tmem_O, tmem_S + k * 8, make_smem_desc(sV + k * 2048, kHalf, 1024), idesc_o, (j | k) ? 1u : 0u);
        tcgen05_commit(bar_ve);
        tcgen05_commit(bar_o);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int q_idx = qt * BM + row;
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(bar_s, (uint32_t)(j & 1));
      tcgen05_fence_after();
      const int kv0 = j * BN;
      const bool masked = (p.causal && kv0 + BN - 1 > qt * BM) || (kv0 + BN > p.S_kv);
      // ---- pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tcgen05_ld_32x32(tmem_S + lane_base + c * 32, r);
        tcgen05_wait_ld();
        if (masked) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kv = kv0 + c * 32 + i;
            const bool ok = kv < p.S_kv && (!p.causal || kv <= q_idx);
            mx = fmaxf(mx, ok ? __uint_as_float(r[i]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      if (j == 0) {
        m_ref = mx;
      } else {
        const float m_new = fmaxf(m_ref, mx);
        const bool need = (m_new - m_ref) * sl2 > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          // S_j ready ⇒ PV_{j-1} retired (commit covers all earlier MMAs), so O is quiescent here
          const float f = need ? ex2((m_ref - m_new) * sl2) : 1.f;
          if (need) m_ref = m_new;
          l *= f;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tcgen05_ld_32x32(tmem_O + lane_base + c * 32, r);
            tcgen05_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            tcgen05_st_32x32(tmem_O + lane_base + c * 32, r);
          }
        }
      }
      // ---- pass 2: P = exp2(s·scale·log2e − m_ref·scale·log2e), packed bf16 over the S columns
      const float mb = m_ref * sl2;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tcgen05_ld_32x32(tmem_S + lane_base + c * 32, r);
        tcgen05_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2(fmaf(__uint_as_float(r[i]), sl2, -mb));
          float p1 = ex2(fmaf(__uint_as_float(r[i + 1]), sl2, -mb));
          if (masked) {
            const int kv = kv0 + c * 32 + i;
            if (!(kv < p.S_kv && (!p.causal || kv <= q_idx))) p0 = 0.f;
            if (!(kv + 1 < p.S_kv && (!p.causal || kv + 1 <= q_idx))) p1 = 0.f;
          }
          l += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
        tcgen05_st_32x16(tmem_S + lane_base + c * 16, pk);
      }
      tcgen05_wait_st();
      tcgen05_fence_before();
      mbar_arrive(bar_p);
    }
    // ---- epilogue: O / l → bf16, LSE
    mbar_wait(bar_o, (uint32_t)((n_kv - 1) & 1));
    tcgen05_fence_after();
    const float inv_l = 1.f / l;
    const bool row_ok = q_idx < p.S_q;
    __nv_bfloat16* orow = out + (long)b * p.o_sb + (long)q_idx * p.o_ss + (long)head * p.o_sh;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tcgen05_ld_32x32(tmem_O + lane_base + c * 32, r);
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
    if (row_ok) p.lse[((long)b * p.H + head) * p.S_q + q_idx] = m_ref * p.scale + __logf(l);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tcgen05_dealloc(tmem_base, 256);
}

// ---- host side -----------------------------------------------------------------------------------------------------
// A [B, S, H, 128] bf16 view with strides (sb, ss, sh, 1) becomes a 2-D TMA tensor whose rows are `ss` apart; batch and
// head offsets fold into the row coordinate when they are multiples of ss, else into the column coordinate.
struct View { const void* ptr; int B, S, H; long sb, ss, sh; };

static CUtensorMap view_tmap(const View& v, Coord& c) {
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
  return make_tmap_bf16_strided(v.ptr, (uint64_t)rows, (uint64_t)cols, (uint64_t)v.ss, 64, 128);
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
  dim3 grid((S_q + BM - 1) / BM, H, B);
  fa_fwd_kernel<<<grid, kThreads, kSmemFwd, st>>>(tq, tk, tv, (__nv_bfloat16*)out, p);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
