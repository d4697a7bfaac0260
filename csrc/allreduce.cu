// One-shot all-reduce over NVLink peer memory for latency-bound messages (decode-time tensor parallelism: a few KB to a
// few MB after every row-parallel GEMV).  Every rank writes its vector into slot[rank] of EVERY peer's symmetric buffer,
// publishes an epoch flag per (source, CTA), and then sums the `world` slots locally — one kernel, one NVLink traversal,
// no ring steps.  Role: the all-reduce of Row-parallel layers / embeddings without sequence parallelism
// (reference layers.py:1040-1043, mappings.py reduce_from_tensor_model_parallel_region) on the inference path.
//
// Graph-safe: the epoch lives in device memory (advanced by the last CTA), so a captured launch replays correctly.
// Payload is double-buffered by epoch parity: a peer can only be in call n+1 after it has seen this rank's flags of call n,
// which this rank publishes before it starts reading call n's slots; call n+2 needs this rank's flags of call n+1.
#include "common.cuh"
#include "kernels.h"

namespace nxd {

constexpr int kArMaxCtas = 64;

template <typename T>
__global__ void __launch_bounds__(512) oneshot_allreduce_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                                const int64_t* __restrict__ peer_bufs,
                                                                const int64_t* __restrict__ peer_flags, long slot_bytes,
                                                                uint32_t* __restrict__ state, int rank, int world, long numel) {
  // state[0] = completed calls (epoch), state[1] = CTA completion counter of the running call
  const uint32_t epoch = ld_acquire_sys(state) + 1u;
  const long half = slot_bytes * world;                       // bytes of one parity half: world slots
  const long base = (long)(epoch & 1u) * half;
  constexpr int VEC = 16 / sizeof(T);
  const long nvec = numel / VEC;
  const long per_cta = (nvec + gridDim.x - 1) / gridDim.x;
  const long v0 = (long)blockIdx.x * per_cta, v1 = min(nvec, v0 + per_cta);
  // ---- push my slice to every rank's slot[rank] (own rank included: keeps the reduce loop uniform)
  for (int p = 0; p < world; ++p) {
    const int dst = (rank + p) % world;
    uint4* d = (uint4*)((uint8_t*)peer_bufs[dst] + base + (long)rank * slot_bytes);
    for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) d[v] = ((const uint4*)x)[v];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world)
    st_release_sys((uint32_t*)peer_flags[threadIdx.x] + rank * kArMaxCtas + blockIdx.x, epoch);
  if ((int)threadIdx.x < world)
    wait_flag_ge((const uint32_t*)peer_flags[rank] + threadIdx.x * kArMaxCtas + blockIdx.x, epoch);
  __syncthreads();
  // ---- reduce the world slots of my slice in fp32
  const uint8_t* mine = (const uint8_t*)peer_bufs[rank] + base;
  for (long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    for (int p = 0; p < world; ++p) {
      const uint4 raw = *(const uint4*)(mine + (long)p * slot_bytes + v * 16);
      const T* e = (const T*)&raw;
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += to_f32<T>(e[j]);
    }
    uint4 o;
    T* oe = (T*)&o;
#pragma unroll
    for (int j = 0; j < VEC; ++j) oe[j] = from_f32<T>(acc[j]);
    ((uint4*)out)[v] = o;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(state + 1, 1u) == gridDim.x - 1) {
      state[1] = 0;
      __threadfence();
      st_release_sys(state, epoch);            // next launch (stream-ordered) sees the advanced epoch
    }
  }
}

void oneshot_allreduce(const void* x, void* out, const int64_t* peer_bufs, const int64_t* peer_flags, long slot_bytes,
                       uint32_t* state, int rank, int world, long numel, int dt, int ctas, cudaStream_t st) {
  if (ctas > kArMaxCtas) ctas = kArMaxCtas;
  if (dt == kBF16)
    oneshot_allreduce_kernel<__nv_bfloat16><<<ctas, 512, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, peer_bufs,
                                                                  peer_flags, slot_bytes, state, rank, world, numel);
  else if (dt == kF32)
    oneshot_allreduce_kernel<float><<<ctas, 512, 0, st>>>((const float*)x, (float*)out, peer_bufs, peer_flags, slot_bytes,
                                                          state, rank, world, numel);
  else
    nxd_throw("oneshot_allreduce: bf16 or fp32 only", __FILE__, __LINE__);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
