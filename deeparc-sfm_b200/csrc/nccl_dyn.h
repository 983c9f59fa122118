// NCCL is bound at run time (dlopen) so that libdeeparc_ba.so has no link-time dependency:
// a single-GPU handle never touches NCCL, and inside a torch process the already loaded
// libnccl.so.2 (torch's bundled copy) is reused instead of mixing two NCCL builds.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace dba {

// ABI-stable subset of nccl.h (NCCL 2.x)
struct NcclUniqueId {
  char internal[128];
};
typedef struct ncclComm* NcclComm;
enum { kNcclSum = 0, kNcclMax = 2, kNcclFloat64 = 8, kNcclUint8 = 1 };

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) return false;
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GetErrorString;
  }
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  return api;
}

}  // namespace dba
