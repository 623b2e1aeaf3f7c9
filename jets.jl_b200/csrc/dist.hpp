// Shared declarations of the multi-GPU translation units (dist.cu: NCCL plumbing and the halo calls on
// caller-owned vectors; dist_op.cu: the distributed operator objects).  Not part of the ABI.
#pragma once
#include <map>
#include "common.hpp"

namespace jets {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

struct Dist {
  Nccl n;
  ncclComm_t comm = nullptr;
  int rank = 0, size = 1;
  bool ready = false;             // rank / size are known (NCCL communicator and/or host bootstrap)
  jets_allgather_fn host_ag = nullptr;   // host bootstrap (jets_dist_init_host): out-of-band all-gather of small records
  void* host_ag_user = nullptr;
  double* dev_gather = nullptr;  // [size] doubles
};
Dist& dist();
void need_dist();
void need_nccl();   // the call moves payload through NCCL: jets_dist_init (not only the host bootstrap) is required
// All-gathers `bytes` bytes per rank (host in, host out[size*bytes]) through the host bootstrap when one was
// given, else through NCCL (synchronises the stream).
void dist_allgather_host(const void* mine, void* all, size_t bytes);
void dist_ops_shutdown();   // dist_op.cu: closes the peer mappings of live distributed operators

#define NCCL_TRY(expr)                                                                      \
  do {                                                                                      \
    int r__ = (expr);                                                                       \
    if (r__ != ncclSuccess)                                                                 \
      JETS_FAIL(JETS_ERR_NCCL, "NCCL error %s at %s:%d", dist().n.GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

}  // namespace jets
