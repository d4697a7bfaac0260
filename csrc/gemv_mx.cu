// Decode-time GEMV on microscaling (OCP MX) weights: y[M<=8, N] = x[M, K] · dequant(W)[N, K]ᵀ (+ residual), with
//   MXFP4: W = e2m1 codes, two per byte (low nibble = even element), K contiguous;  MXFP8: one e4m3 / e5m2 byte per element;
//   one E8M0 scale (2^(s-127)) per 32 consecutive K elements of a row.
// Decode is weight-bandwidth bound, so the win of MX is read traffic: 4.25 bits (fp4) or 8.25 bits (fp8) per weight instead
// of 16.  Nothing is expanded to bf16 in memory: each lane owns one 32-element block per step (16 or 32 bytes of codes + one
// scale byte), decodes in registers (16-entry table in shared memory for e2m1), multiplies with the activation block and applies
// the block scale once to the block's partial sum.  One warp per output row, fp32 accumulation.
// Role: reference quantization_layers.py:626-700 / moe_fused_tkg_mx.py (MX weights at token generation); the prefill-side
// block-scaled tensor-core GEMM (`kind::mxf8f6f4.block_scale`) is a separate kernel family.
#include <cuda_fp8.h>

#include "common.cuh"
#include "kernels.h"

namespace nxd {

namespace {

__device__ __constant__ float kE2M1[16] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f, -0.f, -0.5f, -1.f, -1.5f, -2.f, -3.f, -4.f, -6.f};

NXD_DEVICE float e8m0_to_float(unsigned s) {               // 2^(s-127); s = 0 is the denormal 2^-127, 255 = NaN
  return s == 0u ? __uint_as_float(0x00400000u) : (s == 255u ? __uint_as_float(0x7fc00000u) : __uint_as_float(s << 23));
}

NXD_DEVICE uint4 ld_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// FMT: 0 = e2m1 x2 per byte, 1 = e4m3, 2 = e5m2
template <int M, int FMT>
__global__ void __launch_bounds__(256) gemv_mx_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ w,
                                                      const uint8_t* __restrict__ scale, const __nv_bfloat16* __restrict__ residual,
                                                      __nv_bfloat16* __restrict__ y, int N, int K) {
  __shared__ float lut[16];
  if (threadIdx.x < 16) lut[threadIdx.x] = kE2M1[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  constexpr int kBlockBytes = FMT == 0 ? 16 : 32;
  const int nblk = K / 32;
  const uint8_t* wr = w + (long)n * nblk * kBlockBytes;
  const uint8_t* sr = scale + (long)n * nblk;
  float acc[M];
#pragma unroll
  for (int m = 0; m < M; ++m) acc[m] = 0.f;
  for (int b = lane; b < nblk; b += 32) {
    float v[32];
    if (FMT == 0) {
      const uint4 raw = ld_stream(wr + (long)b * 16);
      const uint8_t* c = reinterpret_cast<const uint8_t*>(&raw);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v[2 * j] = lut[c[j] & 0xF]; v[2 * j + 1] = lut[c[j] >> 4]; }
    } else {
      uint4 raw[2];
      raw[0] = ld_stream(wr + (long)b * 32);
      raw[1] = ld_stream(wr + (long)b * 32 + 16);
      const uint8_t* c = reinterpret_cast<const uint8_t*>(raw);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const __half_raw h = __nv_cvt_fp8_to_halfraw(c[j], FMT == 1 ? __NV_E4M3 : __NV_E5M2);
        v[j] = __half2float(*reinterpret_cast<const __half*>(&h));
      }
    }
    const float s = e8m0_to_float(sr[b]);
#pragma unroll
    for (int m = 0; m < M; ++m) {
      const uint4* xp = reinterpret_cast<const uint4*>(x + (long)m * K + (long)b * 32);
      float part = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 xv = xp[q];
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(x2[i]);
          part = fmaf(f.x, v[q * 8 + 2 * i], part);
          part = fmaf(f.y, v[q * 8 + 2 * i + 1], part);
        }
      }
      acc[m] = fmaf(part, s, acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < M; ++m) {
    const float r = warp_sum(acc[m]);
    if (lane == 0) {
      float o = r;
      if (residual) o += __bfloat162float(residual[(long)m * N + n]);
      y[(long)m * N + n] = __float2bfloat16_rn(o);
    }
  }
}

// Per-slot expert selection for MoE decode: slot s multiplies its own row x[s] with the weights of expert `expert[s]` taken from
// the stacked [E, N, K] MX tensor — only the chosen experts' bytes are read, ids stay on the device (CUDA-graph capturable).
// One warp per (slot, output column); grid = (ceil(N/8), S).
template <int FMT>
__global__ void __launch_bounds__(256) gemv_mx_grouped_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ w,
                                                              const uint8_t* __restrict__ scale, const long* __restrict__ expert,
                                                              __nv_bfloat16* __restrict__ y, int N, int K, int E) {
  __shared__ float lut[16];
  if (threadIdx.x < 16) lut[threadIdx.x] = kE2M1[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp, s = blockIdx.y;
  if (n >= N) return;
  long e = expert[s];
  e = e < 0 ? 0 : (e >= E ? E - 1 : e);
  constexpr int kBlockBytes = FMT == 0 ? 16 : 32;
  const int nblk = K / 32;
  const uint8_t* wr = w + ((long)e * N + n) * nblk * kBlockBytes;
  const uint8_t* sr = scale + ((long)e * N + n) * nblk;
  const __nv_bfloat16* xr = x + (long)s * K;
  float acc = 0.f;
  for (int b = lane; b < nblk; b += 32) {
    float v[32];
    if (FMT == 0) {
      const uint4 raw = ld_stream(wr + (long)b * 16);
      const uint8_t* c = reinterpret_cast<const uint8_t*>(&raw);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v[2 * j] = lut[c[j] & 0xF]; v[2 * j + 1] = lut[c[j] >> 4]; }
    } else {
      uint4 raw[2];
      raw[0] = ld_stream(wr + (long)b * 32);
      raw[1] = ld_stream(wr + (long)b * 32 + 16);
      const uint8_t* c = reinterpret_cast<const uint8_t*>(raw);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const __half_raw h = __nv_cvt_fp8_to_halfraw(c[j], FMT == 1 ? __NV_E4M3 : __NV_E5M2);
        v[j] = __half2float(*reinterpret_cast<const __half*>(&h));
      }
    }
    const uint4* xp = reinterpret_cast<const uint4*>(xr + (long)b * 32);
    float part = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 xv = xp[q];
      const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&xv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(x2[i]);
        part = fmaf(f.x, v[q * 8 + 2 * i], part);
        part = fmaf(f.y, v[q * 8 + 2 * i + 1], part);
      }
    }
    acc = fmaf(part, e8m0_to_float(sr[b]), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) y[(long)s * N + n] = __float2bfloat16_rn(acc);
}

template <int FMT>
void launch(const void* x, const void* w, const void* scale, const void* residual, void* y, int M, int N, int K, cudaStream_t st) {
  const int grid = (N + 7) / 8;
#define NXD_GMX(Mv)                                                                                                     \
  gemv_mx_kernel<Mv, FMT><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const uint8_t*)w, (const uint8_t*)scale,      \
                                                (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, N, K)
  switch (M) {
    case 1: NXD_GMX(1); break;
    case 2: NXD_GMX(2); break;
    case 3: NXD_GMX(3); break;
    case 4: NXD_GMX(4); break;
    case 5: NXD_GMX(5); break;
    case 6: NXD_GMX(6); break;
    case 7: NXD_GMX(7); break;
    case 8: NXD_GMX(8); break;
    default: nxd_throw("gemv_mx: 1 <= M <= 8", __FILE__, __LINE__);
  }
#undef NXD_GMX
}

}  // namespace

void gemv_mx_grouped(const void* x, const void* w, const void* scale, const long* expert, void* y, int S, int N, int K, int E, int fmt,
                     cudaStream_t st) {
  if (K % 32) nxd_throw("gemv_mx_grouped: K must be a multiple of the MX block (32)", __FILE__, __LINE__);
  if (S == 0) return;
  const dim3 grid((N + 7) / 8, S);
#define NXD_GMXG(F)                                                                                                    \
  gemv_mx_grouped_kernel<F><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const uint8_t*)w, (const uint8_t*)scale, expert, \
                                                  (__nv_bfloat16*)y, N, K, E)
  if (fmt == 0) NXD_GMXG(0);
  else if (fmt == 1) NXD_GMXG(1);
  else if (fmt == 2) NXD_GMXG(2);
  else nxd_throw("gemv_mx_grouped: fmt 0 (e2m1), 1 (e4m3) or 2 (e5m2)", __FILE__, __LINE__);
#undef NXD_GMXG
  NXD_CUDA_CHECK(cudaGetLastError());
}

void gemv_mx(const void* x, const void* w, const void* scale, const void* residual, void* y, int M, int N, int K, int fmt,
             cudaStream_t st) {
  if (K % 32) nxd_throw("gemv_mx: K must be a multiple of the MX block (32)", __FILE__, __LINE__);
  if (fmt == 0) launch<0>(x, w, scale, residual, y, M, N, K, st);
  else if (fmt == 1) launch<1>(x, w, scale, residual, y, M, N, K, st);
  else if (fmt == 2) launch<2>(x, w, scale, residual, y, M, N, K, st);
  else nxd_throw("gemv_mx: fmt 0 (e2m1), 1 (e4m3) or 2 (e5m2)", __FILE__, __LINE__);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
