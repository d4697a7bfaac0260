// Decode-time MoE block in ONE launch (role of the reference's `moe_block_tkg` NKI mega-kernel, K8 — modules/moe/
// moe_fused_tkg.py:274-380): for T <= 8 tokens
//     RMSNorm -> router GEMV -> activation / top-k / normalisation -> gate|up GEMVs of the chosen LOCAL experts -> GLU
//     activation -> down GEMV -> affinity-weighted sum
// A decode step is weight-bandwidth bound (every chosen expert's 3·H·I weights are read once) and, as separate launches, launch
// bound: ~12 kernels per layer in the composed path.  Here one persistent cooperative kernel walks the phases with three grid
// barriers; inside a phase the work items (expert, 256-column chunk, K slice) are spread over every warp of the grid so that
// all SMs stream weights, partial sums meet in fp32 scratch through vector `red.global.add`.
//
// Layouts are the framework's (= the reference's): x [T,H] bf16, router weight [E,H], gate|up [E_local, H, 2·I_local] (gate
// columns first), down [E_local, I_local, H].  The output is this rank's PARTIAL sum over its intermediate shard / local
// experts — the caller all-reduces (the "delayed all-reduce" of the reference).  Intermediate values stay fp32 (the composed
// path rounds them to bf16 between its kernels).
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace nxd {

namespace {

constexpr int kTkgThreads = 256;
constexpr int kTkgWarps = kTkgThreads / 32;
constexpr int kTkgMaxT = 8;
constexpr int kTkgMaxK = 8;
constexpr int kTkgMaxE = 256;
constexpr int kTkgMaxActive = kTkgMaxT * kTkgMaxK;       // (token, slot) pairs → at most that many distinct experts
constexpr int kGuSlice = 128;                            // rows of H per gate|up work item
constexpr int kDnSlice = 64;                             // rows of I per down work item (2 per lane)

struct TkgParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* gamma;      // may be null: no norm
  const __nv_bfloat16* router_w;   // [E, H]
  const float* router_bias;        // [E] or null
  const __nv_bfloat16* w_gu;       // [El, H, 2I]
  const __nv_bfloat16* w_dn;       // [El, I, H]
  float* logits;                   // [T, E]   out
  float* gu;                       // [T*K, 2I] scratch (zeroed inside)
  float* yacc;                     // [T, H]   scratch (zeroed inside)
  __nv_bfloat16* out;              // [T, H]
  long* topk_idx;                  // [T, K]   out
  float* topk_w;                   // [T, K]   out
  unsigned* barrier;               // one counter, zero at launch
  int T, H, E, El, e0, I, K;
  float eps;
  int router_act;                  // 0 softmax, 1 sigmoid
  int act_over_topk, normalize, pre_scale, round_logits;
  int act;                         // 0 silu, 1 gelu (erf), 2 gelu (tanh), 3 swiglu: g·σ(α g)·(u+β), 4 relu
  float act_alpha, act_beta, gate_lo, gate_hi, up_lo, up_hi;
};

NXD_DEVICE unsigned ld_acquire_gpu_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
NXD_DEVICE void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
NXD_DEVICE void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
NXD_DEVICE uint4 ld_stream16(const void* p) {            // weights are read exactly once: do not pollute L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
NXD_DEVICE void unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
NXD_DEVICE float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// All CTAs of the (co-resident) grid meet; `target` = barriers passed so far × gridDim.x.
NXD_DEVICE void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_gpu_add(ctr, 1u);
    uint32_t it = 0;
    uint64_t t0 = 0;
    while (ld_acquire_gpu_u32(ctr) < target) {
      __nanosleep(32);
      if ((++it & 0x3fffu) == 0) {                       // a grid that is not co-resident would wait forever: trap after ~4 s
        uint64_t t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        else if (t - t0 > 4000000000ull) __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

NXD_DEVICE float glu_act(const TkgParams& p, float g, float u) {
  g = fminf(fmaxf(g, p.gate_lo), p.gate_hi);
  u = fminf(fmaxf(u, p.up_lo), p.up_hi);
  switch (p.act) {
    case 0: return g / (1.f + __expf(-g)) * u;
    case 1: return 0.5f * g * (1.f + erff(g * 0.70710678118654752f)) * u;
    case 2: return 0.5f * g * (1.f + tanhf(0.7978845608028654f * (g + 0.044715f * g * g * g))) * u;
    case 3: return g / (1.f + __expf(-p.act_alpha * g)) * (u + p.act_beta);
    default: return fmaxf(g, 0.f) * u;
  }
}

struct ValIdx { float v; int i; };
NXD_DEVICE ValIdx vi_better(ValIdx a, ValIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }
NXD_DEVICE ValIdx vi_warp_best(ValIdx x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = vi_better(x, y);
  }
  return x;
}

__global__ void __launch_bounds__(kTkgThreads) moe_block_tkg_kernel(const TkgParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [T][H] normalised tokens
  __shared__ int s_e[kTkgMaxT][kTkgMaxK];
  __shared__ float s_w[kTkgMaxT][kTkgMaxK];
  __shared__ int s_na;
  __shared__ int s_act_e[kTkgMaxActive];                                    // local expert id of active expert a
  __shared__ unsigned s_act_mask[kTkgMaxActive];                            // tokens routed to it
  __shared__ float s_act_w[kTkgMaxActive][kTkgMaxT];
  __shared__ int s_slot[kTkgMaxActive][kTkgMaxT];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = p.T, H = p.H, E = p.E, I = p.I, K = p.K, N2 = 2 * p.I;
  const int gw = blockIdx.x * kTkgWarps + warp, GW = gridDim.x * kTkgWarps;
  const long gtid = (long)blockIdx.x * kTkgThreads + tid, gthreads = (long)gridDim.x * kTkgThreads;

  // ---- phase A: zero the scratch, RMSNorm every token into shared memory (each CTA keeps its own copy) ----------------
  for (long i = gtid; i < (long)T * K * N2; i += gthreads) p.gu[i] = 0.f;
  for (long i = gtid; i < (long)T * H; i += gthreads) p.yacc[i] = 0.f;
  for (int t = warp; t < T; t += kTkgWarps) {
    const __nv_bfloat16* xr = p.x + (long)t * H;
    float rstd = 1.f;
    if (p.gamma != nullptr) {
      float ss = 0.f;
      for (int i = lane * 8; i < H; i += 256) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(xr + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
      }
      rstd = rsqrtf(warp_sum(ss) / (float)H + p.eps);
    }
    for (int i = lane * 8; i < H; i += 256) {
      uint4 raw = *reinterpret_cast<const uint4*>(xr + i);
      if (p.gamma != nullptr) {
        float f[8], g[8];
        unpack8(raw, f);
        unpack8(*reinterpret_cast<const uint4*>(p.gamma + i), g);
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(f[2 * j] * rstd * g[2 * j], f[2 * j + 1] * rstd * g[2 * j + 1]);
      }
      *reinterpret_cast<uint4*>(hs + (long)t * H + i) = raw;
    }
  }
  __syncthreads();

  // ---- phase B: router logits, one warp per expert row -------------------------------------------------------------
  for (int e = gw; e < E; e += GW) {
    float acc[kTkgMaxT];
#pragma unroll
    for (int t = 0; t < kTkgMaxT; ++t) acc[t] = 0.f;
    const __nv_bfloat16* wr = p.router_w + (long)e * H;
    for (int i = lane * 8; i < H; i += 256) {
      float wf[8];
      unpack8(ld_stream16(wr + i), wf);
#pragma unroll
      for (int t = 0; t < kTkgMaxT; ++t) {
        if (t < T) {
          float hf[8];
          unpack8(*reinterpret_cast<const uint4*>(hs + (long)t * H + i), hf);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t] = fmaf(hf[j], wf[j], acc[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < kTkgMaxT; ++t) {
      if (t < T) {
        float v = warp_sum(acc[t]);
        if (lane == 0) {
          if (p.router_bias != nullptr) v += p.router_bias[e];
          p.logits[(long)t * E + e] = p.round_logits ? bf16_round(v) : v;
        }
      }
    }
  }
  grid_barrier(p.barrier, 1u * gridDim.x);

  // ---- phase C: activation, top-k, normalisation — every CTA computes the same routing table ----------------------------
  for (int t = warp; t < T; t += kTkgWarps) {
    float v[kTkgMaxE / 32], a[kTkgMaxE / 32];
    float mx = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < kTkgMaxE / 32; ++j) {
      const int e = lane + 32 * j;
      v[j] = e < E ? __ldcg(p.logits + (long)t * E + e) : -FLT_MAX;
      mx = fmaxf(mx, v[j]);
    }
    if (!p.act_over_topk) {
      if (p.router_act == 0) {
        mx = warp_max(mx);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kTkgMaxE / 32; ++j) { a[j] = (lane + 32 * j) < E ? __expf(v[j] - mx) : 0.f; s += a[j]; }
        s = warp_sum(s);
#pragma unroll
        for (int j = 0; j < kTkgMaxE / 32; ++j) a[j] /= s;
      } else {
#pragma unroll
        for (int j = 0; j < kTkgMaxE / 32; ++j) a[j] = 1.f / (1.f + __expf(-v[j]));
      }
    }
    float sel_w[kTkgMaxK];
    int sel_e[kTkgMaxK];
#pragma unroll
    for (int kk = 0; kk < kTkgMaxK; ++kk) {
      if (kk < K) {
        ValIdx best{-FLT_MAX, 0x7fffffff};
#pragma unroll
        for (int j = 0; j < kTkgMaxE / 32; ++j) {
          const int e = lane + 32 * j;
          if (e < E) best = vi_better(best, ValIdx{p.act_over_topk ? v[j] : a[j], e});
        }
        best = vi_warp_best(best);
        sel_e[kk] = best.i;
        sel_w[kk] = best.v;                                          // logit (act_over_topk) or affinity
#pragma unroll
        for (int j = 0; j < kTkgMaxE / 32; ++j)
          if (lane + 32 * j == best.i) { v[j] = -FLT_MAX; a[j] = -FLT_MAX; }
      }
    }
    if (p.act_over_topk) {
      if (p.router_act == 0) {
        float m2 = -FLT_MAX, s = 0.f;
#pragma unroll
        for (int kk = 0; kk < kTkgMaxK; ++kk) if (kk < K) m2 = fmaxf(m2, sel_w[kk]);
#pragma unroll
        for (int kk = 0; kk < kTkgMaxK; ++kk) if (kk < K) { sel_w[kk] = __expf(sel_w[kk] - m2); s += sel_w[kk]; }
#pragma unroll
        for (int kk = 0; kk < kTkgMaxK; ++kk) if (kk < K) sel_w[kk] /= s;
      } else {
#pragma unroll
        for (int kk = 0; kk < kTkgMaxK; ++kk) if (kk < K) sel_w[kk] = 1.f / (1.f + __expf(-sel_w[kk]));
      }
    }
    // the module hands affinities on in the activation dtype (bf16) and normalises them there
    float s = 0.f;
#pragma unroll
    for (int kk = 0; kk < kTkgMaxK; ++kk) if (kk < K) { sel_w[kk] = bf16_round(sel_w[kk]); s += sel_w[kk]; }
    if (p.normalize) {
      s = fmaxf(bf16_round(s), 1e-9f);
#pragma unroll
      for (int kk = 0; kk < kTkgMaxK; ++kk) if (kk < K) sel_w[kk] = bf16_round(sel_w[kk] / s);
    }
    if (lane == 0) {
#pragma unroll
      for (int kk = 0; kk < kTkgMaxK; ++kk) {
        if (kk < K) {
          s_e[t][kk] = sel_e[kk];
          s_w[t][kk] = sel_w[kk];
          if (blockIdx.x == 0) { p.topk_idx[(long)t * K + kk] = sel_e[kk]; p.topk_w[(long)t * K + kk] = sel_w[kk]; }
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0) {                                                      // ≤ 64 pairs: serial is fine
    int na = 0, nslots = 0;
    for (int t = 0; t < T; ++t) {
      for (int kk = 0; kk < K; ++kk) {
        const int le = s_e[t][kk] - p.e0;
        if (le < 0 || le >= p.El) continue;                           // another EP rank owns this expert
        int a = 0;
        while (a < na && s_act_e[a] != le) ++a;
        if (a == na) {
          s_act_e[a] = le;
          s_act_mask[a] = 0u;
          for (int q = 0; q < kTkgMaxT; ++q) { s_act_w[a][q] = 0.f; s_slot[a][q] = 0; }
          ++na;
        }
        if (s_act_mask[a] >> t & 1u) { s_act_w[a][t] += s_w[t][kk]; continue; }   // the same expert twice for a token
        s_act_mask[a] |= 1u << t;
        s_act_w[a][t] = s_w[t][kk];
        s_slot[a][t] = nslots++;
      }
    }
    s_na = na;
  }
  __syncthreads();
  const int na = s_na;

  // ---- phase D: gate|up partial GEMVs: item = (active expert, 256-column chunk of 2I, 128-row slice of H) ---------------
  {
    const int nchunk = (N2 + 255) / 256, nks = (H + kGuSlice - 1) / kGuSlice;
    const int items = na * nchunk * nks;
    for (int it = gw; it < items; it += GW) {
      const int a = it / (nchunk * nks), r = it % (nchunk * nks), c = r / nks, ks = r % nks;
      const int col = c * 256 + lane * 8;
      const bool live = col < N2;
      const unsigned mask = s_act_mask[a];
      const int k0 = ks * kGuSlice, k1 = min(kGuSlice, H - k0);
      const __nv_bfloat16* wp = p.w_gu + ((long)s_act_e[a] * H + k0) * N2 + (live ? col : 0);
      float acc[kTkgMaxT][8];
#pragma unroll
      for (int t = 0; t < kTkgMaxT; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
      for (int kb = 0; kb < k1; kb += 8) {
        uint4 rows[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)                                   // 8 independent 16-byte loads in flight per lane
          rows[q] = (live && kb + q < k1) ? ld_stream16(wp + (long)(kb + q) * N2) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float wf[8];
          unpack8(rows[q], wf);
          const int k = min(k0 + kb + q, H - 1);
#pragma unroll
          for (int t = 0; t < kTkgMaxT; ++t) {
            if (mask >> t & 1u) {
              const float hv = __bfloat162float(hs[(long)t * H + k]);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(hv, wf[j], acc[t][j]);
            }
          }
        }
      }
      if (live) {
#pragma unroll
        for (int t = 0; t < kTkgMaxT; ++t) {
          if (mask >> t & 1u) {
            float* dst = p.gu + (long)s_slot[a][t] * N2 + col;
            red_add_v4(dst, acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
            red_add_v4(dst + 4, acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
          }
        }
      }
    }
  }
  grid_barrier(p.barrier, 2u * gridDim.x);

  // ---- phase E: GLU activation on the fly + down GEMV: item = (active expert, 256-column chunk of H, 64-row slice of I) --
  {
    const int nchunk = (H + 255) / 256, nks = (I + kDnSlice - 1) / kDnSlice;
    const int items = na * nchunk * nks;
    for (int it = gw; it < items; it += GW) {
      const int a = it / (nchunk * nks), r = it % (nchunk * nks), c = r / nks, ks = r % nks;
      const int col = c * 256 + lane * 8;
      const bool live = col < H;
      const unsigned mask = s_act_mask[a];
      const int k0 = ks * kDnSlice, k1 = min(kDnSlice, I - k0);
      float av[kTkgMaxT][2];                                          // activated inputs of rows k0+lane, k0+32+lane
#pragma unroll
      for (int t = 0; t < kTkgMaxT; ++t) {
        av[t][0] = av[t][1] = 0.f;
        if (mask >> t & 1u) {
          const float w = s_act_w[a][t];
          const float* gsrc = p.gu + (long)s_slot[a][t] * N2;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int k = k0 + q * 32 + lane;
            if (k < I) {
              float g = __ldcg(gsrc + k), u = __ldcg(gsrc + I + k);
              if (p.pre_scale) { g *= w; u *= w; }
              av[t][q] = glu_act(p, g, u) * (p.pre_scale ? 1.f : w);
            }
          }
        }
      }
      const __nv_bfloat16* wp = p.w_dn + ((long)s_act_e[a] * I + k0) * H + (live ? col : 0);
      float acc[kTkgMaxT][8];
#pragma unroll
      for (int t = 0; t < kTkgMaxT; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
      for (int kb = 0; kb < k1; kb += 8) {
        uint4 rows[8];
#pragma unroll
        for (int q = 0; q < 8; ++q)
          rows[q] = (live && kb + q < k1) ? ld_stream16(wp + (long)(kb + q) * H) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float wf[8];
          unpack8(rows[q], wf);
          const int kk = kb + q;                                      // row inside the slice, < 64
#pragma unroll
          for (int t = 0; t < kTkgMaxT; ++t) {
            if (mask >> t & 1u) {
              const float lo = __shfl_sync(0xffffffffu, av[t][0], kk & 31);
              const float hi = __shfl_sync(0xffffffffu, av[t][1], kk & 31);
              const float hv = kk < 32 ? lo : hi;
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(hv, wf[j], acc[t][j]);
            }
          }
        }
      }
      if (live) {
#pragma unroll
        for (int t = 0; t < kTkgMaxT; ++t) {
          if (mask >> t & 1u) {
            float* dst = p.yacc + (long)t * H + col;
            red_add_v4(dst, acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
            red_add_v4(dst + 4, acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
          }
        }
      }
    }
  }
  grid_barrier(p.barrier, 3u * gridDim.x);

  // ---- phase F: fp32 accumulators → bf16 output ------------------------------------------------------------------------
  for (long i = gtid * 8; i < (long)T * H; i += gthreads * 8) {
    uint4 raw;
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(__ldcg(p.yacc + i + 2 * j), __ldcg(p.yacc + i + 2 * j + 1));
    *reinterpret_cast<uint4*>(p.out + i) = raw;
  }
}

}  // namespace

bool moe_block_tkg_supported(int T, int H, int E, int I, int K) {
  return T >= 1 && T <= kTkgMaxT && K >= 1 && K <= kTkgMaxK && K <= E && E <= kTkgMaxE && H % 8 == 0 && I % 8 == 0 &&
         (size_t)T * H * 2 <= 200 * 1024;
}

void moe_block_tkg(const void* x, const void* gamma, const void* router_w, const float* router_bias, const void* w_gu,
                   const void* w_dn, float* logits, float* gu, float* yacc, void* out, long* topk_idx, float* topk_w,
                   unsigned* barrier, int T, int H, int E, int El, int e0, int I, int K, float eps, int router_act,
                   int act_over_topk, int normalize, int pre_scale, int round_logits, int act, float act_alpha, float act_beta,
                   float gate_lo, float gate_hi, float up_lo, float up_hi, bool cooperative, cudaStream_t st) {
  if (!moe_block_tkg_supported(T, H, E, I, K)) nxd_throw("moe_block_tkg: T<=8, K<=8, E<=256, H%8==0, I%8==0", __FILE__, __LINE__);
  TkgParams p;
  p.x = (const __nv_bfloat16*)x; p.gamma = (const __nv_bfloat16*)gamma; p.router_w = (const __nv_bfloat16*)router_w;
  p.router_bias = router_bias; p.w_gu = (const __nv_bfloat16*)w_gu; p.w_dn = (const __nv_bfloat16*)w_dn;
  p.logits = logits; p.gu = gu; p.yacc = yacc; p.out = (__nv_bfloat16*)out; p.topk_idx = topk_idx; p.topk_w = topk_w;
  p.barrier = barrier;
  p.T = T; p.H = H; p.E = E; p.El = El; p.e0 = e0; p.I = I; p.K = K; p.eps = eps; p.router_act = router_act;
  p.act_over_topk = act_over_topk; p.normalize = normalize; p.pre_scale = pre_scale; p.round_logits = round_logits;
  p.act = act; p.act_alpha = act_alpha; p.act_beta = act_beta; p.gate_lo = gate_lo; p.gate_hi = gate_hi; p.up_lo = up_lo;
  p.up_hi = up_hi;
  const size_t smem = (size_t)T * H * sizeof(__nv_bfloat16);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    NXD_CUDA_CHECK(cudaGetDevice(&dev));
    NXD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  NXD_CUDA_CHECK(cudaFuncSetAttribute(moe_block_tkg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  NXD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, moe_block_tkg_kernel, kTkgThreads, smem));
  if (per_sm < 1) nxd_throw("moe_block_tkg: the kernel does not fit on an SM with this T*H", __FILE__, __LINE__);
  const int grid = sms * (per_sm > 2 ? 2 : per_sm);                    // every CTA must be resident: the phases meet at barriers
  NXD_CUDA_CHECK(cudaMemsetAsync(barrier, 0, sizeof(unsigned), st));
  if (cooperative) {
    void* args[] = {(void*)&p};
    NXD_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)moe_block_tkg_kernel, dim3(grid), dim3(kTkgThreads), args, smem, st));
  } else {
    moe_block_tkg_kernel<<<grid, kTkgThreads, smem, st>>>(p);
    NXD_CUDA_CHECK(cudaGetLastError());
  }
}

}  // namespace nxd
