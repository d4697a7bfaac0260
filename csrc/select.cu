// Row-wise selection kernels for on-device sampling (roles of the reference's NKI `cascaded_max` K10, `nkilib.core.topk`
// K11 and the torch_neuronx TopK / Argmax custom calls K12 — operators/argmax.py:129, operators/topk.py:17-21,
// utils/sampling.py:50-60).  A vocab(-shard) row is a one-pass HBM-bound reduction: one CTA per row, 16-byte loads,
// ties resolved to the smallest index (torch.argmax / torch.topk order on equal values is unspecified; the distributed
// wrappers in operators/ rely on "smallest global index wins").
//   row_argmax   x [R, V] → value fp32 [R], index int64 [R] (+ index_offset: global id of this rank's first column)
//   row_topk     x [R, V] → values fp32 [R, k], indices int64 [R, k], sorted descending; k <= 128, V*4 bytes <= 200 KB:
//                the row is staged once in shared memory as fp32 and k block-wide arg-max passes pick the winners.
#include <cfloat>

#include "common.cuh"
#include "kernels.h"

namespace nxd {

namespace {

struct ValIdx { float v; int i; };

NXD_DEVICE ValIdx better(ValIdx a, ValIdx b) {           // larger value, then smaller index
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
NXD_DEVICE ValIdx warp_best(ValIdx x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx y;
    y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
    y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  return x;
}
NXD_DEVICE ValIdx block_best(ValIdx x, ValIdx* smem) {    // smem: >= 32 entries; result valid in all threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  x = warp_best(x);
  __syncthreads();
  if (lane == 0) smem[warp] = x;
  __syncthreads();
  ValIdx r = lane < nw ? smem[lane] : ValIdx{-FLT_MAX, 0x7fffffff};
  r = warp_best(r);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256) row_argmax_kernel(const T* __restrict__ x, float* __restrict__ val, long* __restrict__ idx,
                                                         int V, long row_stride, long index_offset) {
  __shared__ ValIdx red[32];
  const T* row = x + (long)blockIdx.x * row_stride;
  ValIdx best{-FLT_MAX, 0x7fffffff};
  constexpr int N = 16 / sizeof(T);
  const bool vec_ok = ((uintptr_t)row % 16 == 0);
  const int nvec = vec_ok ? V / N : 0;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    Pack16<T> p;
    p.load(row + (long)v * N);
#pragma unroll
    for (int j = 0; j < N; ++j) best = better(best, ValIdx{p.f(j), v * N + j});
  }
  for (int c = nvec * N + threadIdx.x; c < V; c += blockDim.x) best = better(best, ValIdx{to_f32<T>(row[c]), c});
  best = block_best(best, red);
  if (threadIdx.x == 0) { val[blockIdx.x] = best.v; idx[blockIdx.x] = (long)best.i + index_offset; }
}

template <typename T>
__global__ void __launch_bounds__(256) row_topk_kernel(const T* __restrict__ x, float* __restrict__ vals, long* __restrict__ idxs,
                                                       int V, int k, long row_stride, long index_offset) {
  extern __shared__ float srow[];
  __shared__ ValIdx red[32];
  const T* row = x + (long)blockIdx.x * row_stride;
  for (int c = threadIdx.x; c < V; c += blockDim.x) srow[c] = to_f32<T>(row[c]);
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    ValIdx best{-FLT_MAX, 0x7fffffff};
    for (int c = threadIdx.x; c < V; c += blockDim.x) best = better(best, ValIdx{srow[c], c});
    best = block_best(best, red);
    if (threadIdx.x == 0) {
      vals[(long)blockIdx.x * k + j] = best.v;
      idxs[(long)blockIdx.x * k + j] = (long)best.i + index_offset;
      if (best.i < V) srow[best.i] = -FLT_MAX;
    }
    __syncthreads();
  }
}

}  // namespace

void row_argmax(const void* x, float* val, long* idx, int rows, int V, long row_stride, long index_offset, int dt, cudaStream_t st) {
  if (rows == 0) return;
  if (dt == kBF16) row_argmax_kernel<__nv_bfloat16><<<rows, 256, 0, st>>>((const __nv_bfloat16*)x, val, idx, V, row_stride, index_offset);
  else if (dt == kF16) row_argmax_kernel<__half><<<rows, 256, 0, st>>>((const __half*)x, val, idx, V, row_stride, index_offset);
  else row_argmax_kernel<float><<<rows, 256, 0, st>>>((const float*)x, val, idx, V, row_stride, index_offset);
  NXD_CUDA_CHECK(cudaGetLastError());
}

bool row_topk_supported(int V, int k) { return k >= 1 && k <= 128 && (size_t)V * 4 <= 200 * 1024 && k <= V; }

void row_topk(const void* x, float* vals, long* idxs, int rows, int V, int k, long row_stride, long index_offset, int dt,
              cudaStream_t st) {
  if (rows == 0) return;
  if (!row_topk_supported(V, k)) nxd_throw("row_topk: k <= 128 and a row of at most 51200 elements", __FILE__, __LINE__);
  const size_t smem = (size_t)V * 4;
#define NXD_TOPK(T)                                                                                              \
  do {                                                                                                           \
    auto kern = row_topk_kernel<T>;                                                                              \
    if (smem > 48 * 1024) NXD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<rows, 256, smem, st>>>((const T*)x, vals, idxs, V, k, row_stride, index_offset);                      \
  } while (0)
  if (dt == kBF16) NXD_TOPK(__nv_bfloat16);
  else if (dt == kF16) NXD_TOPK(__half);
  else NXD_TOPK(float);
#undef NXD_TOPK
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd

// =====================================================================================================================
// MoE block-metadata build (role of the reference's `find_nonzero_indices` / `indexed_flatten` NKI kernels K9,
// modules/moe/expert_mlps_v2.py:1079-1206): from the routed expert ids [T, k] produce, in ONE launch and without a sort,
//   tokens_per_expert [E], block_to_expert [nb], token_position_to_id [nb * B]  (-1 = padding)
// with every expert's tokens in ascending token order (the stable order of the torch.argsort-based path it replaces).
// One CTA of 1024 threads: smem histogram → per-expert block ranges → chunked stable scatter (warp match + cross-warp prefix).
namespace nxd {
namespace {

constexpr int kMoeThreads = 1024;
constexpr int kMoeMaxExperts = 256;

template <typename IdxT>
__global__ void __launch_bounds__(kMoeThreads) moe_block_metadata_kernel(const IdxT* __restrict__ expert_index, long n, int k, int E,
                                                                         int B, int nb, long* __restrict__ block_to_expert,
                                                                         long* __restrict__ tp2id, long* __restrict__ counts_out) {
  __shared__ int counts[kMoeMaxExperts];
  __shared__ int base[kMoeMaxExperts];
  __shared__ int wc[32][kMoeMaxExperts];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < E; e += kMoeThreads) counts[e] = 0;
  for (long s = tid; s < (long)nb * B; s += kMoeThreads) tp2id[s] = -1;
  __syncthreads();
  for (long i = tid; i < n; i += kMoeThreads) {
    const int e = (int)expert_index[i];
    if (e >= 0 && e < E) atomicAdd(&counts[e], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int blk = 0;
    for (int e = 0; e < E; ++e) {
      const int nblk = (counts[e] + B - 1) / B;
      base[e] = blk * B;                               // first slot of expert e
      for (int b = blk; b < blk + nblk && b < nb; ++b) block_to_expert[b] = e;
      blk += nblk;
    }
    for (int b = blk; b < nb; ++b) block_to_expert[b] = E - 1;        // unused tail blocks (all padding)
  }
  for (int e = tid; e < E; e += kMoeThreads) counts_out[e] = counts[e];
  __syncthreads();
  for (long chunk = 0; chunk < n; chunk += kMoeThreads) {
    for (int j = tid; j < 32 * E; j += kMoeThreads) wc[j / E][j % E] = 0;
    __syncthreads();
    const long i = chunk + tid;
    int e = -1;
    if (i < n) { e = (int)expert_index[i]; if (e < 0 || e >= E) e = -1; }
    const unsigned mask = __match_any_sync(0xffffffffu, e);
    const int rank_in_warp = __popc(mask & ((1u << lane) - 1u));
    if (e >= 0 && rank_in_warp == 0) wc[warp][e] = __popc(mask);
    __syncthreads();
    if (e >= 0) {
      int before = 0;
      for (int w = 0; w < warp; ++w) before += wc[w][e];
      const long slot = (long)base[e] + before + rank_in_warp;
      if (slot < (long)nb * B) tp2id[slot] = i / k;
    }
    __syncthreads();
    for (int ee = tid; ee < E; ee += kMoeThreads) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += wc[w][ee];
      base[ee] += tot;
    }
    __syncthreads();
  }
}

}  // namespace

void moe_block_metadata(const void* expert_index, bool idx64, long n, int k, int E, int B, int nb, long* block_to_expert, long* tp2id,
                        long* counts, cudaStream_t st) {
  if (E > kMoeMaxExperts) nxd_throw("moe_block_metadata: at most 256 experts", __FILE__, __LINE__);
  if (idx64) moe_block_metadata_kernel<long><<<1, kMoeThreads, 0, st>>>((const long*)expert_index, n, k, E, B, nb, block_to_expert, tp2id, counts);
  else moe_block_metadata_kernel<int><<<1, kMoeThreads, 0, st>>>((const int*)expert_index, n, k, E, B, nb, block_to_expert, tp2id, counts);
  NXD_CUDA_CHECK(cudaGetLastError());
}

}  // namespace nxd
