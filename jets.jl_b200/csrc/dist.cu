// Multi-GPU plumbing: one process per GPU: rank bookkeeping, the NCCL communicator (dense-structure
// all-gather / reduce-scatter, bootstrap) and the host bootstrap.  The banded operators' halo traffic does not
// come through here: it moves over peer memory inside the apply kernel (dist_op.cu).
// NCCL is resolved lazily with dlopen so that (a) the library has no link-time NCCL dependency
// and (b) inside a process that already loaded torch's bundled libnccl.so.2 the same copy is
// reused (two NCCL copies in one process would each build their own transport state).
#include <dlfcn.h>
#include "dist.hpp"

namespace jets {
Dist& dist() {
  static Dist d;
  return d;
}
namespace {

void load_nccl() {
  Nccl& n = dist().n;
  if (n.lib) return;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.lib) break;
  }
  JETS_CHECK(n.lib, JETS_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                         \
  n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name));             \
  JETS_CHECK(n.field, JETS_ERR_NCCL, "libnccl is missing symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(ReduceScatter, "ncclReduceScatter");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

int nccl_type(int dt) {
  // the multi-GPU exchange paths are built and measured for the real eltypes (complex vectors would need
  // their element counts doubled for NCCL and for the rank-ordered adds): fail loudly instead of mis-sizing
  JETS_CHECK(!is_cplx(dt), JETS_ERR_UNSUPPORTED, "jets_dist_*: complex eltypes are not implemented on the multi-GPU path");
  return dt == JETS_F32 ? ncclFloat32 : ncclFloat64;
}

__global__ void sum_in_order_kernel(const double* v, int n, double* out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i];
  *out = s;
}

}  // namespace

void need_dist() { JETS_CHECK(dist().ready, JETS_ERR_NCCL, "jets_dist_init() has not been called"); }
void need_nccl() {
  JETS_CHECK(dist().ready && dist().comm, JETS_ERR_NCCL, "this call needs the NCCL communicator: jets_dist_init() has not been called");
}

void dist_allgather_host(const void* mine, void* all, size_t bytes) {
  Dist& d = dist();
  if (d.host_ag) {
    const int rc = d.host_ag(d.host_ag_user, mine, all, (int64_t)bytes);
    JETS_CHECK(rc == 0, JETS_ERR_NCCL, "the host bootstrap's all-gather callback failed with code %d", rc);
    return;
  }
  need_nccl();
  char* dev = nullptr;
  CUDA_TRY(cudaMalloc(&dev, (size_t)(d.size + 1) * bytes));
  try {
    CUDA_TRY(cudaMemcpyAsync(dev + (size_t)d.size * bytes, mine, bytes, cudaMemcpyHostToDevice, ctx().stream));
    NCCL_TRY(d.n.AllGather(dev + (size_t)d.size * bytes, dev, bytes, ncclInt8, d.comm, ctx().stream));
    CUDA_TRY(cudaMemcpyAsync(all, dev, (size_t)d.size * bytes, cudaMemcpyDeviceToHost, ctx().stream));
    CUDA_TRY(cudaStreamSynchronize(ctx().stream));
  } catch (...) { cudaFree(dev); throw; }
  cudaFree(dev);
}
}  // namespace jets

using namespace jets;

extern "C" {

int jets_dist_unique_id(char id[128]) {
  return guard([&] {
    load_nccl();
    ncclUniqueId u;
    NCCL_TRY(dist().n.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
  });
}

int jets_dist_init(int rank, int nranks, const char id[128]) {
  return guard([&] {
    require_ready();
    load_nccl();
    Dist& d = dist();
    JETS_CHECK(!d.comm, JETS_ERR_NCCL, "jets_dist_init called twice");
    JETS_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, JETS_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    JETS_CHECK(!d.ready || (d.rank == rank && d.size == nranks), JETS_ERR_INVALID, "jets_dist_init after jets_dist_init_host with another rank/size");
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    NCCL_TRY(d.n.CommInitRank(&d.comm, nranks, u, rank));
    d.rank = rank;
    d.size = nranks;
    CUDA_TRY(cudaMalloc(&d.dev_gather, (nranks + 2) * sizeof(double)));
    d.ready = true;
  });
}

int jets_dist_init_host(int rank, int nranks, jets_allgather_fn allgather, void* user) {
  return guard([&] {
    require_ready();
    Dist& d = dist();
    JETS_CHECK(allgather, JETS_ERR_INVALID, "null all-gather callback");
    JETS_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, JETS_ERR_INVALID, "bad rank %d of %d", rank, nranks);
    JETS_CHECK(!d.ready || (d.rank == rank && d.size == nranks), JETS_ERR_INVALID, "jets_dist_init_host after jets_dist_init with another rank/size");
    d.host_ag = allgather;
    d.host_ag_user = user;
    d.rank = rank;
    d.size = nranks;
    d.ready = true;
  });
}

int jets_dist_shutdown(void) {
  return guard([&] {
    Dist& d = dist();
    if (!d.ready) return;
    cudaStreamSynchronize(ctx().stream);
    dist_ops_shutdown();
    if (d.comm) d.n.CommDestroy(d.comm);
    if (d.dev_gather) cudaFree(d.dev_gather);
    d.dev_gather = nullptr;
    d.host_ag = nullptr; d.host_ag_user = nullptr;
    d.comm = nullptr; d.ready = false; d.size = 1; d.rank = 0;
  });
}
int jets_dist_rank(void) { return dist().rank; }
int jets_dist_size(void) { return dist().size; }

int jets_dist_sum_scalar(double* inout) {
  return guard([&] {
    require_ready(); need_dist();
    Dist& d = dist();
    Context& c = ctx();
    if (d.host_ag) {   // rank order on the host: bit-stable
      std::vector<double> all(d.size);
      dist_allgather_host(inout, all.data(), sizeof(double));
      double s = 0.0;
      for (double v : all) s += v;
      *inout = s;
      return;
    }
    need_nccl();
    c.host_scratch[40] = *inout;
    double* mine = d.dev_gather + d.size;
    CUDA_TRY(cudaMemcpyAsync(mine, &c.host_scratch[40], sizeof(double), cudaMemcpyHostToDevice, c.stream));
    NCCL_TRY(d.n.AllGather(mine, d.dev_gather, 1, ncclFloat64, d.comm, c.stream));
    sum_in_order_kernel<<<1, 1, 0, c.stream>>>(d.dev_gather, d.size, mine + 1);  // rank order: bit-stable
    count_launch();
    CUDA_TRY(cudaMemcpyAsync(&c.host_scratch[41], mine + 1, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    CUDA_TRY(cudaStreamSynchronize(c.stream));
    *inout = c.host_scratch[41];
  });
}

int jets_dist_allgather(jets_buf shard, jets_buf full) {
  return guard([&] {
    require_ready(); need_nccl();
    Dist& d = dist();
    JETS_CHECK(full->length() == shard->length() * d.size, JETS_ERR_SHAPE, "allgather: full must be nranks x shard");
    NCCL_TRY(d.n.AllGather(shard->ptr(), full->ptr(), (size_t)shard->length(), nccl_type(shard->dtype), d.comm, ctx().stream));
  });
}

int jets_dist_reduce_scatter(jets_buf full, jets_buf shard) {
  return guard([&] {
    require_ready(); need_nccl();
    Dist& d = dist();
    JETS_CHECK(full->length() == shard->length() * d.size, JETS_ERR_SHAPE, "reduce_scatter: full must be nranks x shard");
    NCCL_TRY(d.n.ReduceScatter(full->ptr(), shard->ptr(), (size_t)shard->length(), nccl_type(full->dtype), ncclSum, d.comm, ctx().stream));
  });
}

}  // extern "C"
