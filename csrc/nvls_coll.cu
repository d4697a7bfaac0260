// Stand-alone collectives over an NVLS symmetric region (csrc/symm_vmm.cpp): one kernel each, no NCCL.
//   nvls_allreduce       latency-bound all-reduce (decode-time RowParallelLinear / ParallelEmbedding without sequence
//                        parallel; reference layers.py:1040-1043, mappings.py:196-211): every rank copies its vector into its
//                        OWN slot, one `multimem.red` per CTA signals all ranks at once, then every rank reads the in-switch
//                        sum with `multimem.ld_reduce` (fp32 accumulation) — one local write + one reduced read per element,
//                        versus world peer writes + world local reads in the unicast kernel (allreduce.cu).
//   nvls_all_gather      `multimem.st` of the local shard to all ranks + cross-rank CTA barrier.
//   nvls_reduce_scatter  `multimem.ld_reduce` of this rank's chunk of a symmetric input.
// All three are CUDA-graph capturable: the call counter (epoch) lives in device memory and is advanced by the last CTA.
// Region layout is chosen by the caller (ops/nvls.py): `flag_off` = kNvlsCollMaxCtas u32 barrier counters, payload
// double-buffered by epoch parity.  With `mc_base == nullptr` the kernels use unicast peer accesses (same protocol).
#include "common.cuh"
#include "kernels.h"

namespace nxd {

constexpr int kNvlsCollMaxCtas = 128;

NXD_DEVICE void mm_red_add_release_u32(uint32_t* mc, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc), "r"(v) : "memory");
}
NXD_DEVICE void mm_st_v4(void* mc, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <typename T> NXD_DEVICE uint4 mm_ld_reduce(const void* mc);
template <> NXD_DEVICE uint4 mm_ld_reduce<__nv_bfloat16>(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}
template <> NXD_DEVICE uint4 mm_ld_reduce<float>(const void* mc) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mc) : "memory");
  return v;
}

struct NvlsRegion {
  const int64_t* peer_bases;   // [world] unicast VAs
  uint8_t* mc_base;            // multicast VA or nullptr
  uint8_t* local_base;
  long flag_off;               // kNvlsCollMaxCtas u32 counters
  long data_off;               // payload: 2 parity halves of `half_bytes`
  long half_bytes;
  int rank, world;
};

// Cross-rank barrier between CTA `blockIdx.x` of every rank: counter[cta] += 1 on all ranks, wait until it shows `world`
// arrivals for this CTA's n-th barrier (n is kept per CTA in state[2 + cta], so calls may use different grid sizes as long
// as every rank uses the same one).  Caller has fenced its payload writes (fence.sys) and synchronised the CTA.
NXD_DEVICE void cta_barrier_all_ranks(const NvlsRegion& r, uint32_t* state) {
  if (r.mc_base != nullptr) {
    if (threadIdx.x == 0) mm_red_add_release_u32((uint32_t*)(r.mc_base + r.flag_off) + blockIdx.x, 1u);
  } else if ((int)threadIdx.x < r.world) {
    red_add_release_sys((uint32_t*)((uint8_t*)r.peer_bases[threadIdx.x] + r.flag_off) + blockIdx.x, 1u);
  }
  if (threadIdx.x == 0) {
    const uint32_t n = state[2 + blockIdx.x] + 1u;
    state[2 + blockIdx.x] = n;
    wait_flag_ge((const uint32_t*)(r.local_base + r.flag_off) + blockIdx.x, n * (uint32_t)r.world);
  }
  __syncthreads();
}

NXD_DEVICE void finish_call(uint32_t* state, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(state + 1, 1u) == gridDim.x - 1) {
      state[1] = 0;
      __threadfence();
      st_release_sys(state, epoch);
    }
  }
}

// out = sum over ranks of x (+ residual).  numel * sizeof(T) multiple of 16 and <= half_bytes.
template <typename T>
__global__ void __launch_bounds__(512) nvls_allreduce_kernel(const T* __restrict__ x, const T* __restrict__ residual,
                                                             T* __restrict__ out, NvlsRegion r, uint32_t* __restrict__ state,
                                                             long numel) {
  const uint32_t epoch = ld_acquire_sys(state) + 1u;
  const long base = r.data_off + (long)(epoch & 1u) * r.half_bytes;
  constexpr int VEC = 16 / sizeof(T);
  const long nvec = numel / VEC;
  const long per_cta = (nvec + gridDim.x - 1) / gridDim.x;
  const long v0 = (long)blockIdx.x * per_cta, v1 = min(nvec, v0 + per_cta);
  uint4* mine = (uint4*)(r.local_base + base);
  for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) mine[v] = ((const uint4*)x)[v];
  __threadfence_system();
  __syncthreads();
  cta_barrier_all_ranks(r, state);
  for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    uint4 o;
    if (r.mc_base != nullptr) {
      o = mm_ld_reduce<T>(r.mc_base + base + v * 16);
    } else {
      float acc[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
      for (int p = 0; p < r.world; ++p) {
        uint4 raw;
        const void* src = (const uint8_t*)r.peer_bases[p] + base + v * 16;
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(src) : "memory");
        const T* e = (const T*)&raw;
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += to_f32<T>(e[j]);
      }
      T* oe = (T*)&o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) oe[j] = from_f32<T>(acc[j]);
    }
    if (residual != nullptr) {
      const uint4 rr = ((const uint4*)residual)[v];
      const T* re = (const T*)&rr;
      T* oe = (T*)&o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) oe[j] = from_f32<T>(to_f32<T>(oe[j]) + to_f32<T>(re[j]));
    }
    ((uint4*)out)[v] = o;
  }
  finish_call(state, epoch);
}

// Every rank contributes `bytes` (multiple of 16); afterwards payload[parity][p*bytes .. ) holds rank p's data on all ranks.
// `out` (optional) receives a private copy of the gathered buffer.
__global__ void __launch_bounds__(512) nvls_all_gather_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ out,
                                                              NvlsRegion r, uint32_t* __restrict__ state, long bytes) {
  const uint32_t epoch = ld_acquire_sys(state) + 1u;
  const long base = r.data_off + (long)(epoch & 1u) * r.half_bytes;
  const long nvec = bytes / 16;
  const long per_cta = (nvec + gridDim.x - 1) / gridDim.x;
  const long v0 = (long)blockIdx.x * per_cta, v1 = min(nvec, v0 + per_cta);
  const long dst_off = base + (long)r.rank * bytes;
  constexpr int U = 4;
  for (long b = v0 + threadIdx.x; b < v1; b += (long)U * blockDim.x) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const long i = b + (long)u * blockDim.x; if (i < v1) v[u] = ((const uint4*)x)[i]; }
    if (r.mc_base != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u) { const long i = b + (long)u * blockDim.x; if (i < v1) mm_st_v4(r.mc_base + dst_off + i * 16, v[u]); }
    } else {
      for (int p = 0; p < r.world; ++p) {
        uint4* d = (uint4*)((uint8_t*)r.peer_bases[p] + dst_off);
#pragma unroll
        for (int u = 0; u < U; ++u) { const long i = b + (long)u * blockDim.x; if (i < v1) d[i] = v[u]; }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  cta_barrier_all_ranks(r, state);
  if (out != nullptr) {
    // slice [v0, v1) of every rank's contribution has landed (CTA i of every rank moved the same slice)
    for (int p = 0; p < r.world; ++p) {
      const uint4* s = (const uint4*)(r.local_base + base + (long)p * bytes);
      uint4* d = (uint4*)(out + (long)p * bytes);
      for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        uint4 t;
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(s + v) : "memory");
        d[v] = t;
      }
    }
  }
  finish_call(state, epoch);
}

// x = [world * chunk] elements on every rank; out[chunk] = sum over ranks of x[rank*chunk .. (rank+1)*chunk).
template <typename T>
__global__ void __launch_bounds__(512) nvls_reduce_scatter_kernel(const T* __restrict__ x, T* __restrict__ out, NvlsRegion r,
                                                                  uint32_t* __restrict__ state, long chunk) {
  const uint32_t epoch = ld_acquire_sys(state) + 1u;
  const long base = r.data_off + (long)(epoch & 1u) * r.half_bytes;
  constexpr int VEC = 16 / sizeof(T);
  const long cvec = chunk / VEC;
  const long per_cta = (cvec + gridDim.x - 1) / gridDim.x;
  const long v0 = (long)blockIdx.x * per_cta, v1 = min(cvec, v0 + per_cta);
  // copy slice [v0, v1) of EVERY chunk into the own slot: CTA i of rank q later reads slice i of chunk q from all ranks
  for (int p = 0; p < r.world; ++p) {
    const uint4* s = (const uint4*)x + (long)p * cvec;
    uint4* d = (uint4*)(r.local_base + base) + (long)p * cvec;
    for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) d[v] = s[v];
  }
  __threadfence_system();
  __syncthreads();
  cta_barrier_all_ranks(r, state);
  const long my = base + (long)r.rank * chunk * (long)sizeof(T);
  for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    uint4 o;
    if (r.mc_base != nullptr) {
      o = mm_ld_reduce<T>(r.mc_base + my + v * 16);
    } else {
      float acc[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
      for (int p = 0; p < r.world; ++p) {
        uint4 raw;
        const void* src = (const uint8_t*)r.peer_bases[p] + my + v * 16;
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(src) : "memory");
        const T* e = (const T*)&raw;
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] += to_f32<T>(e[j]);
      }
      T* oe = (T*)&o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) oe[j] = from_f32<T>(acc[j]);
    }
    ((uint4*)out)[v] = o;
  }
  finish_call(state, epoch);
}

static NvlsRegion make_region(const int64_t* peer_bases, int64_t mc_base, int64_t local_base, long flag_off, long data_off,
                              long half_bytes, int rank, int world) {
  NvlsRegion r;
  r.peer_bases = peer_bases; r.mc_base = (uint8_t*)mc_base; r.local_base = (uint8_t*)local_base; r.flag_off = flag_off;
  r.data_off = data_off; r.half_bytes = half_bytes; r.rank = rank; r.world = world;
  return r;
}

void nvls_allreduce(const void* x, const void* residual, void* out, const int64_t* peer_bases, int64_t mc_base, int64_t local_base,
                    long flag_off, long data_off, long half_bytes, uint32_t* state, int rank, int world, long numel, int dt,
                    int ctas, cudaStream_t st) {
  if (ctas > kNvlsCollMaxCtas) ctas = kNvlsCollMaxCtas;
  const NvlsRegion r = make_region(peer_bases, mc_base, local_base, flag_off, data_off, half_bytes, rank, world);
  const long bytes = numel * (dt == kF32 ? 4 : 2);
  if (bytes % 16 || bytes > half_bytes) nxd_throw("nvls_allreduce: size must be a multiple of 16 bytes and fit the slot", __FILE__, __LINE__);
  if (dt == kBF16)
    nvls_allreduce_kernel<__nv_bfloat16><<<ctas, 512, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)residual,
                                                               (__nv_bfloat16*)out, r, state, numel);
  else if (dt == kF32)
    nvls_allreduce_kernel<float><<<ctas, 512, 0, st>>>((const float*)x, (const float*)residual, (float*)out, r, state, numel);
  else
    nxd_throw("nvls_allreduce: bf16 or fp32 only", __FILE__, __LINE__);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void nvls_all_gather(const void* x, void* out, const int64_t* peer_bases, int64_t mc_base, int64_t local_base, long flag_off,
                     long data_off, long half_bytes, uint32_t* state, int rank, int world, long bytes, int ctas, cudaStream_t st) {
  if (ctas > kNvlsCollMaxCtas) ctas = kNvlsCollMaxCtas;
  if (bytes % 16 || bytes * world > half_bytes) nxd_throw("nvls_all_gather: bad size", __FILE__, __LINE__);
  const NvlsRegion r = make_region(peer_bases, mc_base, local_base, flag_off, data_off, half_bytes, rank, world);
  nvls_all_gather_kernel<<<ctas, 512, 0, st>>>((const uint8_t*)x, (uint8_t*)out, r, state, bytes);
  NXD_CUDA_CHECK(cudaGetLastError());
}

void nvls_reduce_scatter(const void* x, void* out, const int64_t* peer_bases, int64_t mc_base, int64_t local_base, long flag_off,
                         long data_off, long half_bytes, uint32_t* state, int rank, int world, long chunk_numel, int dt, int ctas,
                         cudaStream_t st) {
  if (ctas > kNvlsCollMaxCtas) ctas = kNvlsCollMaxCtas;
  const long bytes = chunk_numel * (dt == kF32 ? 4 : 2);
  if (bytes % 16 || bytes * world > half_bytes) nxd_throw("nvls_reduce_scatter: bad size", __FILE__, __LINE__);
  const NvlsRegion r = make_region(peer_bases, mc_base, local_base, flag_off, data_off, half_bytes, rank, world);
  if (dt == kBF16)
    nvls_reduce_scatter_kernel<__nv_bfloat16><<<ctas, 512, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, r, state, chunk_numel);
  else if (dt == kF32)
    nvls_reduce_scatter_kernel<float><<<ctas, 512, 0, st>>>((const float*)x, (float*)out, r, state, chunk_numel);
  else
    nxd_throw("nvls_reduce_scatter: bf16 or fp32 only", __FILE__, __LINE__);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
