// Symmetric (peer-mapped) device memory: cudaMalloc + CUDA IPC.  Every rank allocates a payload region and
// a flag region, exports IPC handles (exchanged by python over the gloo control group) and opens all peers'
// handles; the resulting device-pointer tables are what the fused kernels dereference over NVLink.
#include <cuda_runtime.h>

#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "symm.h"

namespace nxd {

namespace {
struct Region {
  void* payload = nullptr;
  void* flags = nullptr;
  size_t nbytes = 0;
  size_t nflags = 0;
  std::vector<void*> opened;  // peer mappings to close
};
std::mutex g_mu;
std::map<int64_t, Region> g_regions;
int64_t g_next = 1;

void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace

SymmAlloc symm_alloc(size_t nbytes, size_t nflags) {
  Region r;
  r.nbytes = nbytes;
  r.nflags = nflags;
  check(cudaMalloc(&r.payload, nbytes), "cudaMalloc(payload)");
  check(cudaMalloc(&r.flags, nflags * sizeof(uint32_t)), "cudaMalloc(flags)");
  check(cudaMemset(r.flags, 0, nflags * sizeof(uint32_t)), "cudaMemset(flags)");
  check(cudaDeviceSynchronize(), "sync");
  cudaIpcMemHandle_t hp, hf;
  check(cudaIpcGetMemHandle(&hp, r.payload), "cudaIpcGetMemHandle(payload)");
  check(cudaIpcGetMemHandle(&hf, r.flags), "cudaIpcGetMemHandle(flags)");
  SymmAlloc out;
  out.payload_handle.assign((const char*)&hp, sizeof(hp));
  out.flags_handle.assign((const char*)&hf, sizeof(hf));
  std::lock_guard<std::mutex> lk(g_mu);
  out.id = g_next++;
  g_regions[out.id] = r;
  return out;
}

SymmPtrs symm_open(int64_t id, int rank, const std::vector<std::string>& payload_handles,
                   const std::vector<std::string>& flags_handles) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_regions.find(id);
  if (it == g_regions.end()) throw std::runtime_error("symm_open: bad handle");
  Region& r = it->second;
  SymmPtrs out;
  const int world = (int)payload_handles.size();
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      out.payload.push_back((int64_t)r.payload);
      out.flags.push_back((int64_t)r.flags);
      continue;
    }
    cudaIpcMemHandle_t hp, hf;
    std::memcpy(&hp, payload_handles[p].data(), sizeof(hp));
    std::memcpy(&hf, flags_handles[p].data(), sizeof(hf));
    void *pp = nullptr, *pf = nullptr;
    check(cudaIpcOpenMemHandle(&pp, hp, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(payload)");
    check(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle(flags)");
    r.opened.push_back(pp);
    r.opened.push_back(pf);
    out.payload.push_back((int64_t)pp);
    out.flags.push_back((int64_t)pf);
  }
  return out;
}

void symm_free(int64_t id) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_regions.find(id);
  if (it == g_regions.end()) return;
  for (void* p : it->second.opened) cudaIpcCloseMemHandle(p);
  cudaFree(it->second.payload);
  cudaFree(it->second.flags);
  g_regions.erase(it);
}

void* symm_local_payload(int64_t id, size_t* nbytes) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_regions.find(id);
  if (it == g_regions.end()) throw std::runtime_error("symm_local_payload: bad handle");
  if (nbytes) *nbytes = it->second.nbytes;
  return it->second.payload;
}

}  // namespace nxd
