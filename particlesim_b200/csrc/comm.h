// comm.h — NCCL, loaded at run time.  libpsim_b200.so does not link libnccl: a single-GPU host never needs it, and
// a multi-GPU host already has one (the system's libnccl.so.2, or the one its framework loaded).  psim_comm_init
// resolves the handful of entry points below with dlopen / dlsym; the types come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace psim {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  bool ok = false;
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  if (api.handle) return api;
  // a copy that is already in the process (e.g. the one torch bundles) wins over the system's
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    const char* env = getenv("PSIM_NCCL_LIB");
    if (env) api.handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  }
  for (const char* nm : names) {
    if (api.handle) break;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
  }
  if (!api.handle) return api;
  auto sym = [&](const char* s) { return dlsym(api.handle, s); };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GetErrorString && api.AllGather &&
           api.AllReduce && api.Broadcast && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
  return api;
}

}  // namespace psim
