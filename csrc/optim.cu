// Multi-tensor optimizer kernels: one launch covers a whole list of tensors through a chunk table
// passed by value in kernel parameters (no per-tensor launches, no host sync).
//   multi_tensor_sq_norm : Σ‖t‖² → atomicAdd into a device fp32 scalar   (grads.py:41-189)
//   multi_tensor_scale   : t *= *scale (device scalar; clip coefficient) (grads.py:238-256)
//   fused_adamw          : fp32 master/moment update + optional bf16 model copy
//                          (utils/adamw_fp32_optim_params.py:91-155)
#include "common.cuh"
#include "kernels.h"

namespace nxd {

constexpr int kMaxTensors = 48;
constexpr int kMaxBlocks = 640;
constexpr long kChunk = 65536;  // elements per CTA-chunk

template <int DEPTH> struct ChunkTable {
  void* ptr[DEPTH][kMaxTensors];
  long numel[kMaxTensors];
  unsigned char block_tensor[kMaxBlocks];
  int block_chunk[kMaxBlocks];
};

template <typename T>
__global__ void __launch_bounds__(256) sq_norm_kernel(ChunkTable<1> tab, float* __restrict__ out) {
  __shared__ float red[32];
  const int t = tab.block_tensor[blockIdx.x];
  const long base = (long)tab.block_chunk[blockIdx.x] * kChunk;
  const long n = min(tab.numel[t] - base, kChunk);
  const T* p = (const T*)tab.ptr[0][t] + base;
  float s = 0.f;
  constexpr int N = Pack16<T>::N;
  if (((uintptr_t)p % 16) == 0) {
    const long nv = n / N;
    for (long v = threadIdx.x; v < nv; v += blockDim.x) {
      Pack16<T> k; k.load(p + v * N);
#pragma unroll
      for (int j = 0; j < N; ++j) { const float f = k.f(j); s += f * f; }
    }
    for (long i = nv * N + threadIdx.x; i < n; i += blockDim.x) { const float f = to_f32<T>(p[i]); s += f * f; }
  } else {
    for (long i = threadIdx.x; i < n; i += blockDim.x) { const float f = to_f32<T>(p[i]); s += f * f; }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(ChunkTable<1> tab, const float* __restrict__ scale) {
  const int t = tab.block_tensor[blockIdx.x];
  const long base = (long)tab.block_chunk[blockIdx.x] * kChunk;
  const long n = min(tab.numel[t] - base, kChunk);
  T* p = (T*)tab.ptr[0][t] + base;
  const float sc = *scale;
  constexpr int N = Pack16<T>::N;
  if (((uintptr_t)p % 16) == 0) {
    const long nv = n / N;
    for (long v = threadIdx.x; v < nv; v += blockDim.x) {
      Pack16<T> k; k.load(p + v * N);
#pragma unroll
      for (int j = 0; j < N; ++j) k.set(j, k.f(j) * sc);
      k.store(p + v * N);
    }
    for (long i = nv * N + threadIdx.x; i < n; i += blockDim.x) p[i] = from_f32<T>(to_f32<T>(p[i]) * sc);
  } else {
    for (long i = threadIdx.x; i < n; i += blockDim.x) p[i] = from_f32<T>(to_f32<T>(p[i]) * sc);
  }
}

template <typename G, typename L, bool HAS_LOW>
__global__ void __launch_bounds__(256) adamw_kernel(ChunkTable<5> tab, float step_size, float beta1, float beta2,
                                                    float eps, float decay_pre, float decay_post, float vscale,
                                                    const float* __restrict__ grad_scale) {
  // One update rule covers both AdamW forms:
  //   p = (p * decay_pre - step_size * m / (sqrt(v * vscale) + eps)) * decay_post
  // torch form : step_size = lr/bc1, vscale = 1/bc2, decay_pre = 1 - lr*wd, decay_post = 1
  // HF form    : step_size = lr*sqrt(bc2)/bc1, vscale = 1, decay_pre = 1, decay_post = 1 - lr*wd
  //              (reference utils/adamw_fp32_optim_params.py:130-155: eps added before bias correction, decay applied last)
  const int t = tab.block_tensor[blockIdx.x];
  const long base = (long)tab.block_chunk[blockIdx.x] * kChunk;
  const long n = min(tab.numel[t] - base, kChunk);
  float* p = (float*)tab.ptr[0][t] + base;
  const G* g = (const G*)tab.ptr[1][t] + base;
  float* m = (float*)tab.ptr[2][t] + base;
  float* v = (float*)tab.ptr[3][t] + base;
  L* low = HAS_LOW ? (L*)tab.ptr[4][t] + base : nullptr;
  const float gs = grad_scale ? *grad_scale : 1.f;
  // 4 elements / thread / iteration: fp32 state moves as 16-byte packets
  const bool aligned = ((uintptr_t)p % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                       ((uintptr_t)g % (4 * sizeof(G)) == 0) && (!HAS_LOW || (uintptr_t)low % (4 * sizeof(L)) == 0);
  long i0 = 0;
  if (aligned) {
    const long nv = n / 4;
    for (long q = threadIdx.x; q < nv; q += blockDim.x) {
      float4 pp = *reinterpret_cast<float4*>(p + q * 4);
      float4 mm = *reinterpret_cast<float4*>(m + q * 4);
      float4 vv = *reinterpret_cast<float4*>(v + q * 4);
      float gg[4];
      if constexpr (sizeof(G) == 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(g + q * 4);
        gg[0] = t4.x; gg[1] = t4.y; gg[2] = t4.z; gg[3] = t4.w;
      } else {
        const uint2 raw = *reinterpret_cast<const uint2*>(g + q * 4);
        const G* h = reinterpret_cast<const G*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) gg[j] = to_f32<G>(h[j]);
      }
      float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = gg[j] * gs;
        ma[j] = beta1 * ma[j] + (1.f - beta1) * gr;
        va[j] = beta2 * va[j] + (1.f - beta2) * gr * gr;
        const float denom = sqrtf(va[j] * vscale) + eps;
        pa[j] = (pa[j] * decay_pre - step_size * ma[j] / denom) * decay_post;
      }
      *reinterpret_cast<float4*>(p + q * 4) = pp;
      *reinterpret_cast<float4*>(m + q * 4) = mm;
      *reinterpret_cast<float4*>(v + q * 4) = vv;
      if constexpr (HAS_LOW) {
        if constexpr (sizeof(L) == 4) {
          *reinterpret_cast<float4*>(low + q * 4) = pp;
        } else {
          uint2 raw; L* h = reinterpret_cast<L*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = from_f32<L>(pa[j]);
          *reinterpret_cast<uint2*>(low + q * 4) = raw;
        }
      }
    }
    i0 = nv * 4;
  }
  for (long i = i0 + threadIdx.x; i < n; i += blockDim.x) {
    const float gr = to_f32<G>(g[i]) * gs;
    const float mi = beta1 * m[i] + (1.f - beta1) * gr;
    const float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
    const float pi = (p[i] * decay_pre - step_size * mi / (sqrtf(vi * vscale) + eps)) * decay_post;
    m[i] = mi; v[i] = vi; p[i] = pi;
    if constexpr (HAS_LOW) low[i] = from_f32<L>(pi);
  }
}

// simple table builder: tensors are split so that no table exceeds the limits
template <int DEPTH, typename Launch>
static void run_tables(const std::vector<TensorRef>* lists, Launch&& launch) {
  ChunkTable<DEPTH> tab;
  int nt = 0, nb = 0;
  auto flush = [&]() { if (nb > 0) launch(tab, nb); nt = 0; nb = 0; };
  for (size_t t = 0; t < lists[0].size(); ++t) {
    const long numel = lists[0][t].numel;
    if (numel == 0) continue;
    const long chunks = (numel + kChunk - 1) / kChunk;
    long c = 0;
    while (c < chunks) {
      if (nt == kMaxTensors || nb == kMaxBlocks) flush();
      // register (the remaining part of) this tensor in the table
      const int slot = nt++;
      const long elem_off = c * kChunk;
      for (int d = 0; d < DEPTH; ++d) tab.ptr[d][slot] = nullptr;
      tab.numel[slot] = numel - elem_off;
      // pointer offsets are applied per list by the caller-provided element sizes → store base; offset via chunk index
      for (int d = 0; d < DEPTH; ++d)
        tab.ptr[d][slot] = lists[d].empty() ? nullptr : lists[d][t].ptr;
      // since we store the *base* pointer, chunk indices stay absolute and numel stays the full size
      tab.numel[slot] = numel;
      while (c < chunks && nb < kMaxBlocks) {
        tab.block_tensor[nb] = (unsigned char)slot;
        tab.block_chunk[nb] = (int)c;
        ++nb; ++c;
      }
    }
  }
  flush();
}

void multi_tensor_sq_norm(const std::vector<TensorRef>& ts, int dt, float* out, cudaStream_t st) {
  const std::vector<TensorRef> lists[1] = {ts};
  run_tables<1>(lists, [&](const ChunkTable<1>& tab, int nb) {
    switch (dt) {
      case kF32: sq_norm_kernel<float><<<nb, 256, 0, st>>>(tab, out); break;
      case kBF16: sq_norm_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(tab, out); break;
      case kF16: sq_norm_kernel<__half><<<nb, 256, 0, st>>>(tab, out); break;
      default: nxd_throw("bad dtype", __FILE__, __LINE__);
    }
  });
  NXD_CUDA_CHECK(cudaGetLastError());
}

void multi_tensor_scale(const std::vector<TensorRef>& ts, int dt, const float* scale, cudaStream_t st) {
  const std::vector<TensorRef> lists[1] = {ts};
  run_tables<1>(lists, [&](const ChunkTable<1>& tab, int nb) {
    switch (dt) {
      case kF32: scale_kernel<float><<<nb, 256, 0, st>>>(tab, scale); break;
      case kBF16: scale_kernel<__nv_bfloat16><<<nb, 256, 0, st>>>(tab, scale); break;
      case kF16: scale_kernel<__half><<<nb, 256, 0, st>>>(tab, scale); break;
      default: nxd_throw("bad dtype", __FILE__, __LINE__);
    }
  });
  NXD_CUDA_CHECK(cudaGetLastError());
}

void fused_adamw(const std::vector<TensorRef>& p, const std::vector<TensorRef>& g, const std::vector<TensorRef>& m,
                 const std::vector<TensorRef>& v, const std::vector<TensorRef>& lowp, int gdt, int ldt, float lr,
                 float beta1, float beta2, float eps, float wd, float bc1, float bc2, const float* grad_scale,
                 int hf_form, cudaStream_t st) {
  const std::vector<TensorRef> lists[5] = {p, g, m, v, lowp};
  const bool has_low = !lowp.empty();
  const float step_size = hf_form ? lr * sqrtf(bc2) / bc1 : lr / bc1;
  const float vscale = hf_form ? 1.f : 1.f / bc2;
  const float decay_pre = hf_form ? 1.f : 1.f - lr * wd;
  const float decay_post = hf_form ? 1.f - lr * wd : 1.f;
  run_tables<5>(lists, [&](const ChunkTable<5>& tab, int nb) {
#define LAUNCH(G, L, HL) adamw_kernel<G, L, HL><<<nb, 256, 0, st>>>(tab, step_size, beta1, beta2, eps, decay_pre, decay_post, vscale, grad_scale)
    if (gdt == kF32) {
      if (!has_low) LAUNCH(float, float, false);
      else if (ldt == kBF16) LAUNCH(float, __nv_bfloat16, true);
      else if (ldt == kF16) LAUNCH(float, __half, true);
      else LAUNCH(float, float, true);
    } else if (gdt == kBF16) {
      if (!has_low) LAUNCH(__nv_bfloat16, float, false);
      else if (ldt == kBF16) LAUNCH(__nv_bfloat16, __nv_bfloat16, true);
      else if (ldt == kF16) LAUNCH(__nv_bfloat16, __half, true);
      else LAUNCH(__nv_bfloat16, float, true);
    } else {
      if (!has_low) LAUNCH(__half, float, false);
      else if (ldt == kF16) LAUNCH(__half, __half, true);
      else if (ldt == kBF16) LAUNCH(__half, __nv_bfloat16, true);
      else LAUNCH(__half, float, true);
    }
#undef LAUNCH
  });
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
