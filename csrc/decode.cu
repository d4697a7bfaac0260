// Token-generation (decode) kernels for sm_100a — both are HBM-bandwidth problems, so they are plain CUDA tuned for wide
// coalesced loads rather than tensor-core kernels:
//   decode_attention : one query token per sequence against the KV cache (flash-decoding: split over the cache length,
//                      GQA group processed together so K/V are read once per kv head), + a combine pass over the splits.
//                      Role of the reference's token-generation attention (modules/attention, models' TKG path).
//   gemv_bf16        : y[M≤8, N] = x[M, K] · W[N, K]ᵀ (+ optional residual) for batch-1..8 decode; every weight byte is read once
//                      with 16-byte loads, fp32 accumulation.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace nxd {

namespace {

constexpr int HD = 128;
constexpr int kDecThreads = 128;

// 16 lanes share one kv position (8 dims = 16 bytes each); a warp covers 2 positions per step, the CTA 8.
template <int G>
__global__ void __launch_bounds__(kDecThreads) decode_attn_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc, const __nv_bfloat16* __restrict__ vc,
    const long* __restrict__ positions, float* __restrict__ part_o, float* __restrict__ part_ml, int B, int H, int Hkv, int L,
    long k_sb, long k_ss, long k_sh, long v_sb, long v_ss, long v_sh, long q_sb, long q_sh, float scale_log2, int splits) {
  const int kvh = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 4, l16 = lane & 15;               // position slot within the warp step, 8-dim slice
  const int n_valid = min((int)positions[b] + 1, L);
  const int per = (n_valid + splits - 1) / splits;
  const int s0 = split * per, s1 = min(n_valid, s0 + per);

  auto load8 = [](const __nv_bfloat16* p, float (&f)[8], float mul) {
    const uint4 a = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x * mul; f[2 * i + 1] = t.y * mul; }
  };
  float qf[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) load8(q + b * q_sb + (long)(kvh * G + g) * q_sh + l16 * 8, qf[g], scale_log2);
  float m[G], l[G], o[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[g][i] = 0.f;
  }
  const __nv_bfloat16* kb = kc + b * k_sb + (long)kvh * k_sh + l16 * 8;
  const __nv_bfloat16* vb = vc + b * v_sb + (long)kvh * v_sh + l16 * 8;
  for (int sbase = s0 + warp * 2; sbase < s1; sbase += 8) {      // warp-uniform trip count (full-mask shuffles inside)
    const int s = sbase + sub;
    const bool live = s < s1;
    float kf[8], vf[8];
    if (live) {
      load8(kb + (long)s * k_ss, kf, 1.f);
      load8(vb + (long)s * v_ss, vf, 1.f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { kf[i] = 0.f; vf[i] = 0.f; }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(qf[g][i], kf[i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 8);            // score·log2e of (position s, head g) in all 16 lanes
      if (live) {
        const float mn = fmaxf(m[g], d);
        const float corr = exp2f(m[g] - mn), p = exp2f(d - mn);
        m[g] = mn;
        l[g] = l[g] * corr + p;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[g][i] = fmaf(o[g][i], corr, p * vf[i]);
      }
    }
  }
  // merge the 2 position slots of the warp (lanes l16, l16+16), then the 4 warps through smem
  __shared__ float sm_o[4][G][HD];
  __shared__ float sm_m[4][G], sm_l[4][G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m[g], 16), l2 = __shfl_xor_sync(0xffffffffu, l[g], 16);
    const float mn = fmaxf(m[g], m2);
    const float c1 = mn == -INFINITY ? 0.f : exp2f(m[g] - mn), c2 = mn == -INFINITY ? 0.f : exp2f(m2 - mn);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float o2 = __shfl_xor_sync(0xffffffffu, o[g][i], 16);
      o[g][i] = o[g][i] * c1 + o2 * c2;
    }
    l[g] = l[g] * c1 + l2 * c2;
    m[g] = mn;
    if (sub == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sm_o[warp][g][l16 * 8 + i] = o[g][i];
      if (l16 == 0) { sm_m[warp][g] = m[g]; sm_l[warp][g] = l[g]; }
    }
  }
  __syncthreads();
  // thread t = dim; loop over heads of the group
  const int d = threadIdx.x;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float mn = fmaxf(fmaxf(sm_m[0][g], sm_m[1][g]), fmaxf(sm_m[2][g], sm_m[3][g]));
    float acc = 0.f, ls = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float c = mn == -INFINITY ? 0.f : exp2f(sm_m[w][g] - mn);
      acc += sm_o[w][g][d] * c;
      ls += sm_l[w][g] * c;
    }
    const long idx = (((long)b * H + kvh * G + g) * splits + split);
    part_o[idx * HD + d] = acc;
    if (d == 0) { part_ml[idx * 2] = mn; part_ml[idx * 2 + 1] = ls; }
  }
}

__global__ void __launch_bounds__(HD) decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                                            __nv_bfloat16* __restrict__ out, int splits, long o_sb, long o_sh,
                                                            int H) {
  const int bh = blockIdx.x, d = threadIdx.x;
  float mn = -INFINITY;
  for (int s = 0; s < splits; ++s) mn = fmaxf(mn, part_ml[((long)bh * splits + s) * 2]);
  float acc = 0.f, ls = 0.f;
  for (int s = 0; s < splits; ++s) {
    const long idx = (long)bh * splits + s;
    const float c = mn == -INFINITY ? 0.f : exp2f(part_ml[idx * 2] - mn);
    acc += part_o[idx * HD + d] * c;
    ls += part_ml[idx * 2 + 1] * c;
  }
  const int b = bh / H, h = bh % H;
  out[b * o_sb + h * o_sh + d] = __float2bfloat16_rn(ls > 0.f ? acc / ls : 0.f);
}

// Same merge over the splits, but the result stays un-normalised: o (relative to the row maximum), the maximum (natural-log
// units) and the row sum — what a rank contributes to distributed flash-decoding, where the final normalisation happens after
// the cross-rank log-sum-exp merge (modules/attention/flash_decode.py).
__global__ void __launch_bounds__(HD) decode_combine_partial_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                                                    float* __restrict__ fin_o, float* __restrict__ fin_ml, int splits) {
  const int bh = blockIdx.x, d = threadIdx.x;
  float mn = -INFINITY;
  for (int s = 0; s < splits; ++s) mn = fmaxf(mn, part_ml[((long)bh * splits + s) * 2]);
  float acc = 0.f, ls = 0.f;
  for (int s = 0; s < splits; ++s) {
    const long idx = (long)bh * splits + s;
    const float c = mn == -INFINITY ? 0.f : exp2f(part_ml[idx * 2] - mn);
    acc += part_o[idx * HD + d] * c;
    ls += part_ml[idx * 2 + 1] * c;
  }
  fin_o[(long)bh * HD + d] = acc;
  if (d == 0) { fin_ml[bh * 2] = mn * 0.6931471805599453f; fin_ml[bh * 2 + 1] = ls; }
}

// y[m, n] = Σ_k x[m, k] · W[n, k]; one warp per output column n (8 columns per 256-thread CTA), M ≤ 8 rows of x kept hot in L1.
template <int M>
__global__ void __launch_bounds__(256) gemv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                   const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ y,
                                                   int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  float acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.f;
  const __nv_bfloat16* wr = w + (long)n * K;
#pragma unroll 4
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wr + k));       // weights stream through once
    const __nv_bfloat162* w2 = reinterpret_cast<const __nv_bfloat162*>(&wv);
    float wf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(w2[i]); wf[2 * i] = f.x; wf[2 * i + 1] = f.y; }
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const uint4 xv = *reinterpret_cast<const uint4*>(x + (long)m * K + k);
      const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(x2[i]);
        acc[m] = fmaf(f.x, wf[2 * i], acc[m]);
        acc[m] = fmaf(f.y, wf[2 * i + 1], acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const float v = warp_sum(acc[m]);
    if (lane == 0) {
      float r = v;
      if (residual) r += __bfloat162float(residual[(long)m * N + n]);
      y[(long)m * N + n] = __float2bfloat16_rn(r);
    }
  }
}

// Decode-time RoPE + KV-cache append in ONE launch (was ~25 tiny elementwise / index_put kernels per layer in the token-
// generation graph): q [B,H,D] rotated into q_out, k rotated and v copied into the cache row positions[b].  One warp per head
// row; cos/sin rows gathered by position.  HF rotate-half convention (csrc/elementwise.cu rope_kernel).
__global__ void __launch_bounds__(128) decode_rope_kv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                            const __nv_bfloat16* __restrict__ v, const long* __restrict__ positions,
                                                            const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                            __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ kc,
                                                            __nv_bfloat16* __restrict__ vc, int B, int H, int Hkv, int D, int L,
                                                            long q_sb, long q_sh, long k_sb, long k_sh, long v_sb, long v_sh,
                                                            long c_sb, long c_ss, long c_sh) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int rows = B * (H + 2 * Hkv);
  if (warp >= rows) return;
  const int b = warp / (H + 2 * Hkv), r = warp % (H + 2 * Hkv);
  long pos = positions[b];
  pos = pos < 0 ? 0 : (pos >= L ? L - 1 : pos);
  const int half = D / 2;
  if (r < H + Hkv) {                                   // q or k head: rotate
    const bool is_q = r < H;
    const int h = is_q ? r : r - H;
    const __nv_bfloat16* src = is_q ? q + b * q_sb + (long)h * q_sh : k + b * k_sb + (long)h * k_sh;
    __nv_bfloat16* dst = is_q ? q_out + ((long)b * H + h) * D : kc + b * c_sb + pos * c_ss + (long)h * c_sh;
    const float* cr = cos_t + pos * half;
    const float* sr = sin_t + pos * half;
    for (int c = lane; c < half; c += 32) {
      const float a = __bfloat162float(src[c]), bb = __bfloat162float(src[half + c]);
      const float cs = cr[c], sn = sr[c];
      dst[c] = __float2bfloat16_rn(a * cs - bb * sn);
      dst[half + c] = __float2bfloat16_rn(bb * cs + a * sn);
    }
  } else {                                             // v head: copy
    const int h = r - H - Hkv;
    const __nv_bfloat16* src = v + b * v_sb + (long)h * v_sh;
    __nv_bfloat16* dst = vc + b * c_sb + pos * c_ss + (long)h * c_sh;
    for (int c = lane; c < D; c += 32) dst[c] = src[c];
  }
}

}  // namespace

void decode_rope_kv(const void* q, const void* k, const void* v, const long* positions, const float* cos_t, const float* sin_t,
                    void* q_out, void* kc, void* vc, int B, int H, int Hkv, int D, int L, long q_sb, long q_sh, long k_sb, long k_sh,
                    long v_sb, long v_sh, long c_sb, long c_ss, long c_sh, cudaStream_t st) {
  const int rows = B * (H + 2 * Hkv);
  const int grid = (rows * 32 + 127) / 128;
  decode_rope_kv_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, positions,
                                              cos_t, sin_t, (__nv_bfloat16*)q_out, (__nv_bfloat16*)kc, (__nv_bfloat16*)vc, B, H, Hkv, D, L,
                                              q_sb, q_sh, k_sb, k_sh, v_sb, v_sh, c_sb, c_ss, c_sh);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void decode_attention(const void* q, const void* k, const void* v, const long* positions, void* out, float* part_o,
                      float* part_ml, int B, int H, int Hkv, int L, const long* ks, const long* vs, long q_sb, long q_sh,
                      long o_sb, long o_sh, float scale, int splits, cudaStream_t st, float* fin_o, float* fin_ml) {
  const int G = H / Hkv;
  dim3 grid(Hkv, B, splits);
  const float sl2 = scale * 1.4426950408889634f;
#define NXD_DEC(Gv)                                                                                                  \
  decode_attn_kernel<Gv><<<grid, kDecThreads, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k,             \
      (const __nv_bfloat16*)v, positions, part_o, part_ml, B, H, Hkv, L, ks[0], ks[1], ks[2], vs[0], vs[1], vs[2], q_sb, \
      q_sh, sl2, splits)
  switch (G) {
    case 1: NXD_DEC(1); break;
    case 2: NXD_DEC(2); break;
    case 4: NXD_DEC(4); break;
    case 8: NXD_DEC(8); break;
    default: nxd_throw("decode_attention: unsupported GQA group size", __FILE__, __LINE__);
  }
#undef NXD_DEC
  if (fin_o != nullptr) decode_combine_partial_kernel<<<B * H, HD, 0, st>>>(part_o, part_ml, fin_o, fin_ml, splits);
  else decode_combine_kernel<<<B * H, HD, 0, st>>>(part_o, part_ml, (__nv_bfloat16*)out, splits, o_sb, o_sh, H);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void gemv_bf16(const void* x, const void* w, const void* residual, void* y, int M, int N, int K, cudaStream_t st) {
  const int grid = (N + 7) / 8;
#define NXD_GEMV(Mv)                                                                                                \
  gemv_kernel<Mv><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)residual, \
                                        (__nv_bfloat16*)y, N, K)
  switch (M) {
    case 1: NXD_GEMV(1); break;
    case 2: NXD_GEMV(2); break;
    case 3: NXD_GEMV(3); break;
    case 4: NXD_GEMV(4); break;
    case 5: NXD_GEMV(5); break;
    case 6: NXD_GEMV(6); break;
    case 7: NXD_GEMV(7); break;
    case 8: NXD_GEMV(8); break;
    default: nxd_throw("gemv_bf16: M must be 1..8", __FILE__, __LINE__);
  }
#undef NXD_GEMV
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
