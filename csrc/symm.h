#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace nxd {
struct SymmAlloc { int64_t id; std::string payload_handle, flags_handle; };
struct SymmPtrs { std::vector<int64_t> payload, flags; };
SymmAlloc symm_alloc(size_t nbytes, size_t nflags);
SymmPtrs symm_open(int64_t id, int rank, const std::vector<std::string>& payload_handles,
                   const std::vector<std::string>& flags_handles);
void symm_free(int64_t id);
void* symm_local_payload(int64_t id, size_t* nbytes);
}  // namespace nxd

// ---- v2: VMM allocations + NVLS multicast (symm_vmm.cpp) ----
namespace nxd {
struct VmmBegin { int64_t id = 0; std::string sock_name; bool multicast_supported = false; size_t size = 0; };
struct VmmPtrs { std::vector<int64_t> peer; int64_t multicast = 0; size_t size = 0; };
VmmBegin vmm_begin(size_t nbytes, int rank, int world, bool want_multicast);
std::string vmm_send(int64_t id, const std::vector<std::string>& sock_names, bool use_multicast);
std::string vmm_recv(int64_t id);
std::string vmm_bind(int64_t id, bool everyone_ok);
VmmPtrs vmm_ptrs(int64_t id, bool multicast_everywhere);
void* vmm_local(int64_t id, size_t* nbytes);
void vmm_free(int64_t id);
bool is_vmm_id(int64_t id);
}  // namespace nxd
