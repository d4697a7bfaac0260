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
