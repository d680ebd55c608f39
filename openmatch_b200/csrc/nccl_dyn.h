// Run-time binding of the handful of NCCL entry points the sharded search uses.  The library is opened with
// dlopen("libnccl.so.2") the first time a communicator is created, so libopenmatch_b200.so itself has no link-time
// NCCL dependency (it loads on a CPU-only box for the symbol tests) and, inside a PyTorch process, resolves to the
// very libnccl torch already mapped.  Prototypes restated from the public NCCL 2.x API (nccl.h); enum values are
// ABI-stable across 2.x.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace om {

struct NcclUid {
  char internal[128];
};
enum { kNcclInt8 = 0, kNcclInt32 = 2, kNcclFloat32 = 7 };
enum { kNcclSum = 0, kNcclMax = 2 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  const char* why = nullptr;  // load failure
};

static inline NcclApi& nccl_api() {
  static NcclApi api;
  if (api.handle || api.why) return api;
  const char* names[] = {getenv("OPENMATCH_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.why = "libnccl.so.2 not found (import torch first, or set OPENMATCH_B200_NCCL to its path)";
    return api;
  }
  auto sym = [&](const char* s) { return dlsym(api.handle, s); };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.AllGather ||
      !api.GetErrorString) {
    api.why = "libnccl.so.2 lacks a required symbol";
    api.handle = nullptr;
  }
  return api;
}

}  // namespace om
