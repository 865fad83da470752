// NCCL plumbing shared by the native distributed solvers (capi_nccl.cu, capi_pcg2.cu).
//
// NCCL is not linked: its entry points are taken with dlopen(RTLD_NOLOAD)/dlsym from the libnccl.so.2 that is
// already in the process (the one torch loaded), so the default library has no NCCL dependency and two NCCL
// versions can never meet in one process.  The communicator is this library's own (ncclCommInitRank with an id
// that the Python side broadcasts through torch.distributed); it belongs to the plan and dies with it.
#pragma once
#include "capi_common.cuh"

#include <dlfcn.h>

#include <mutex>
#include <vector>

#if defined(__has_include) && !defined(SKB_NO_NCCL)
#if __has_include(<nccl.h>)
#include <nccl.h>
#define SKB_HAVE_NCCL_H 1
#endif
#endif

#if defined(SKB_HAVE_NCCL_H)
namespace skb {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string why;
};
NcclApi& nccl();   // capi_nccl.cu

#define SKB_NCCL(call)                                                                                  \
  do {                                                                                                  \
    ncclResult_t _r = (call);                                                                           \
    if (_r != ncclSuccess) return skb::fail(SKB_ECUDA, std::string(#call) + ": " + skb::nccl().GetErrorString(_r)); \
  } while (0)

struct Halo {
  int peer;
  int64_t ns, nr;
  const int32_t *sidx, *ridx;
  double *sbuf, *rbuf;
};

struct Pcg2State;                 // capi_pcg2.cu: work vectors and captured graphs of the single-reduction PCG
void pcg2_destroy(Pcg2State* s);  // capi_pcg2.cu

struct DistNative {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  std::vector<Halo> halo;
  Pcg2State* pcg2 = nullptr;
};

// the plan's native distributed state (skb_plan::dist), or nullptr before skb_nccl_init
inline DistNative* state_of(skb_plan* pl) { return pl ? static_cast<DistNative*>(pl->dist) : nullptr; }

int halo_exchange(DistNative& d, double* v, cudaStream_t st);            // capi_nccl.cu
int all_reduce(DistNative& d, double* buf, size_t count, cudaStream_t st);

}  // namespace skb
#endif  // SKB_HAVE_NCCL_H
