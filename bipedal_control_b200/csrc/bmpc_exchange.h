// Multi-GPU policy exchange behind the C ABI (bmpc_exchange_*): one ncclAllGather of the newest policy slab per MPC tick, on its own stream,
// overlapped with the next tick.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch already loaded in a torch.distributed
// process, or the system library), so libbmpc.so has no link-time dependency on it and single-GPU users never load it.
//
// The few NCCL declarations needed are restated here (nccl.h 2.27 / 2.28: ncclUniqueId is 128 opaque bytes, ncclComm_t / ncclWindow_t are
// opaque pointers, ncclFloat64 = 8; ncclConfig_t as of 2.28.0, selected only if ncclGetVersion() >= 22800).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <climits>
#include <cstdlib>
#include <cstddef>
#include <stdexcept>
#include <string>

namespace bmpc {

struct NcclApi {
  struct UniqueId { char internal[128]; };
  struct ConfigV22800 {   // nccl.h 2.28: typedef struct ncclConfig_v22800 { ... } ncclConfig_t
    size_t size; unsigned int magic; unsigned int version;
    int blocking; int cgaClusterSize; int minCTAs; int maxCTAs; const char* netName; int splitShare; int trafficClass; const char* commName;
    int collnetEnable; int CTAPolicy; int shrinkShare; int nvlsCTAs; int nChannelsPerNetPeer; int nvlinkCentricSched;
  };
  typedef void* Comm; typedef void* Window;
  static constexpr int kFloat64 = 8, kWinCollSymmetric = 0x01, kCtaPolicyZero = 0x02;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*CommInitRankConfig)(Comm*, int, UniqueId, int, ConfigV22800*) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, Comm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*MemAlloc)(void**, size_t) = nullptr;
  int (*MemFree)(void*) = nullptr;
  int (*CommWindowRegister)(Comm, void*, size_t, Window*, int) = nullptr;
  int (*CommWindowDeregister)(Comm, Window) = nullptr;
  void* lib = nullptr;
  int version = 0;

  template <class F> void bind(F& f, const char* name, bool required) {
    f = reinterpret_cast<F>(dlsym(lib, name));
    if (!f && required) throw std::runtime_error(std::string("[bmpc] libnccl.so.2 lacks ") + name);
  }
  void load() {
    if (lib) return;
    // a copy that is already in the process (e.g. the one PyTorch ships and loaded) must be the one we use: two different libnccl.so.2 in one
    // process do not work.  Otherwise BMPC_NCCL_LIBRARY (full path) or the system library.  A process that also uses PyTorch has to import torch
    // BEFORE the first bmpc_exchange_* call for the same reason.
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) { const char* path = getenv("BMPC_NCCL_LIBRARY"); if (path && path[0]) lib = dlopen(path, RTLD_NOW | RTLD_LOCAL); }
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::runtime_error(std::string("[bmpc] cannot load libnccl.so.2 (needed only for bmpc_exchange_*): ") + dlerror());
    bind(GetVersion, "ncclGetVersion", true); bind(GetUniqueId, "ncclGetUniqueId", true); bind(CommInitRank, "ncclCommInitRank", true);
    bind(CommInitRankConfig, "ncclCommInitRankConfig", false); bind(CommDestroy, "ncclCommDestroy", true); bind(AllGather, "ncclAllGather", true);
    bind(GetErrorString, "ncclGetErrorString", true); bind(MemAlloc, "ncclMemAlloc", false); bind(MemFree, "ncclMemFree", false);
    bind(CommWindowRegister, "ncclCommWindowRegister", false); bind(CommWindowDeregister, "ncclCommWindowDeregister", false);
    if (GetVersion(&version) != 0) version = 0;
  }
  void check(int rc, const char* what) const { if (rc != 0) throw std::runtime_error(std::string("[bmpc] NCCL: ") + what + ": " + (GetErrorString ? GetErrorString(rc) : "error")); }
  ConfigV22800 default_config() const {
    ConfigV22800 c{};
    c.size = sizeof(ConfigV22800); c.magic = 0xcafebeef; c.version = (unsigned)version;
    c.blocking = INT_MIN; c.cgaClusterSize = INT_MIN; c.minCTAs = INT_MIN; c.maxCTAs = INT_MIN; c.netName = nullptr; c.splitShare = INT_MIN; c.trafficClass = INT_MIN;
    c.commName = nullptr; c.collnetEnable = INT_MIN; c.CTAPolicy = INT_MIN; c.shrinkShare = INT_MIN; c.nvlsCTAs = INT_MIN; c.nChannelsPerNetPeer = INT_MIN; c.nvlinkCentricSched = INT_MIN;
    return c;
  }
};

inline NcclApi& nccl_api() { static NcclApi api; return api; }

}  // namespace bmpc
