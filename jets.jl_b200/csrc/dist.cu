// Multi-GPU plumbing: one process per GPU, NCCL over NVLink 5 / NVSwitch.
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

// x[0:n] += y[0:n]
template <typename T>
__global__ void add_inplace_kernel(T* __restrict__ x, const T* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = x[i] + y[i];
}
void add_inplace(int dt, void* x, const void* y, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  JETS_CHECK(!is_cplx(dt), JETS_ERR_UNSUPPORTED, "jets_dist_*: complex eltypes are not implemented on the multi-GPU path");
  int64_t g = (n + 1023) / 1024;
  const int64_t cap = (int64_t)ctx().sm_count * 16;
  if (g > cap) g = cap;
  if (dt == JETS_F32) add_inplace_kernel<float><<<(unsigned)g, 256, 0, s>>>((float*)x, (const float*)y, n);
  else add_inplace_kernel<double><<<(unsigned)g, 256, 0, s>>>((double*)x, (const double*)y, n);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

__global__ void sum_in_order_kernel(const double* v, int n, double* out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += v[i];
  *out = s;
}

int64_t blocks_len(jets_buf x, int first, int n) { return x->blk_off[first + n] - x->blk_off[first]; }

const void* alloc_base(jets_buf x) { return x->st->alloc ? x->st->alloc : (const void*)x->st->data; }
Dist::Peer* peer_of(jets_buf x) {
  auto it = dist().peers.find(alloc_base(x));
  return it == dist().peers.end() ? nullptr : &it->second;
}
// Remote twin of a local address inside a registered allocation (all ranks share one layout).
const char* remote(const void* peer_base, jets_buf x, const char* local_ptr) {
  return reinterpret_cast<const char*>(peer_base) + (local_ptr - reinterpret_cast<const char*>(alloc_base(x)));
}
void ensure_copy_streams() {
  Dist& d = dist();
  if (d.copy[0]) return;
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaStreamCreateWithFlags(&d.copy[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&d.ev_copy[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&d.ev_begin, cudaEventDisableTiming));
  CUDA_TRY(cudaMalloc(&d.dev_flag, 2 * sizeof(float)));
  CUDA_TRY(cudaMemset(d.dev_flag, 0, 2 * sizeof(float)));
}
// Every rank has reached this point of its stream (and finished everything before it) once the
// all-reduce completes anywhere: the cross-process fence around peer-memory copies.
void stream_barrier() {
  Dist& d = dist();
  NCCL_TRY(d.n.AllReduce(d.dev_flag, d.dev_flag + 1, 1, ncclFloat32, ncclSum, d.comm, ctx().stream));
}
// Copy-engine pulls from the neighbours, ordered after everything on the context stream.
void pull_begin(char* dst_prev, const char* src_prev, size_t n_prev, char* dst_next, const char* src_next, size_t n_next) {
  Dist& d = dist();
  ensure_copy_streams();
  stream_barrier();
  CUDA_TRY(cudaEventRecord(d.ev_begin, ctx().stream));
  if (n_prev) {
    CUDA_TRY(cudaStreamWaitEvent(d.copy[0], d.ev_begin, 0));
    CUDA_TRY(cudaMemcpyAsync(dst_prev, src_prev, n_prev, cudaMemcpyDeviceToDevice, d.copy[0]));
  }
  if (n_next) {
    CUDA_TRY(cudaStreamWaitEvent(d.copy[1], d.ev_begin, 0));
    CUDA_TRY(cudaMemcpyAsync(dst_next, src_next, n_next, cudaMemcpyDeviceToDevice, d.copy[1]));
  }
  CUDA_TRY(cudaEventRecord(d.ev_copy[0], d.copy[0]));
  CUDA_TRY(cudaEventRecord(d.ev_copy[1], d.copy[1]));
}
void pull_end() {
  Dist& d = dist();
  CUDA_TRY(cudaStreamWaitEvent(ctx().stream, d.ev_copy[0], 0));
  CUDA_TRY(cudaStreamWaitEvent(ctx().stream, d.ev_copy[1], 0));
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
    if (d.halo_tmp) cudaFree(d.halo_tmp);
    for (auto& kv : d.peers) {
      if (kv.second.prev) cudaIpcCloseMemHandle(kv.second.prev);
      if (kv.second.next) cudaIpcCloseMemHandle(kv.second.next);
    }
    d.peers.clear();
    d.comm = nullptr; d.ready = false; d.size = 1; d.rank = 0;
    d.halo_tmp = nullptr; d.halo_tmp_bytes = 0;
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

int jets_dist_halo_exchange(jets_buf x, int32_t nlo, jets_buf lo, int32_t nhi, jets_buf hi) {
  return guard([&] {
    require_ready(); need_nccl();
    Dist& d = dist();
    Context& c = ctx();
    const int nb = x->nblocks();
    const int ty = nccl_type(x->dtype);
    const bool has_prev = d.rank > 0, has_next = d.rank + 1 < d.size;
    NCCL_TRY(d.n.GroupStart());
    // my last nlo blocks are the next rank's `lo`; my first nhi blocks are the previous rank's `hi`
    if (has_next && nlo > 0) NCCL_TRY(d.n.Send(x->block_ptr(nb - nlo), (size_t)blocks_len(x, nb - nlo, nlo), ty, d.rank + 1, d.comm, c.stream));
    if (has_prev && nhi > 0) NCCL_TRY(d.n.Send(x->block_ptr(0), (size_t)blocks_len(x, 0, nhi), ty, d.rank - 1, d.comm, c.stream));
    if (has_prev && nlo > 0 && lo) NCCL_TRY(d.n.Recv(lo->ptr(), (size_t)lo->length(), ty, d.rank - 1, d.comm, c.stream));
    if (has_next && nhi > 0 && hi) NCCL_TRY(d.n.Recv(hi->ptr(), (size_t)hi->length(), ty, d.rank + 1, d.comm, c.stream));
    NCCL_TRY(d.n.GroupEnd());
  });
}

namespace {
struct HaloStage { char* from_prev; char* from_next; int64_t n_from_prev, n_from_next; };
// lo = my partial contribution to the previous rank's last nlo blocks; hi = to the next rank's
// first nhi blocks.  I receive the mirror images into staging memory.
HaloStage halo_reduce_xfer(jets_buf x, int32_t nlo, jets_buf lo, int32_t nhi, jets_buf hi) {
  Dist& d = dist();
  Context& c = ctx();
  const int nb = x->nblocks();
  const int ty = nccl_type(x->dtype);
  const size_t es = dsize(x->dtype);
  const bool has_prev = d.rank > 0, has_next = d.rank + 1 < d.size;
  HaloStage h;
  h.n_from_prev = (has_prev && nhi > 0) ? blocks_len(x, 0, nhi) : 0;
  h.n_from_next = (has_next && nlo > 0) ? blocks_len(x, nb - nlo, nlo) : 0;
  const size_t need = (size_t)(h.n_from_prev + h.n_from_next) * es + 512;
  if (need > d.halo_tmp_bytes) {
    CUDA_TRY(cudaDeviceSynchronize());
    if (d.halo_tmp) cudaFree(d.halo_tmp);
    CUDA_TRY(cudaMalloc(&d.halo_tmp, need));
    d.halo_tmp_bytes = need;
  }
  h.from_prev = d.halo_tmp;
  h.from_next = d.halo_tmp + (((size_t)h.n_from_prev * es + 255) & ~(size_t)255);
  NCCL_TRY(d.n.GroupStart());
  if (has_prev && nlo > 0 && lo) NCCL_TRY(d.n.Send(lo->ptr(), (size_t)lo->length(), ty, d.rank - 1, d.comm, c.stream));
  if (has_next && nhi > 0 && hi) NCCL_TRY(d.n.Send(hi->ptr(), (size_t)hi->length(), ty, d.rank + 1, d.comm, c.stream));
  if (h.n_from_prev) NCCL_TRY(d.n.Recv(h.from_prev, (size_t)h.n_from_prev, ty, d.rank - 1, d.comm, c.stream));
  if (h.n_from_next) NCCL_TRY(d.n.Recv(h.from_next, (size_t)h.n_from_next, ty, d.rank + 1, d.comm, c.stream));
  NCCL_TRY(d.n.GroupEnd());
  return h;
}
void halo_reduce_add(jets_buf x, int32_t nlo, int32_t nhi) {   // previous rank first: deterministic
  Dist& d = dist();
  const int nb = x->nblocks();
  const size_t es = dsize(x->dtype);
  const bool has_prev = d.rank > 0, has_next = d.rank + 1 < d.size;
  const int64_t n_from_prev = (has_prev && nhi > 0) ? blocks_len(x, 0, nhi) : 0;
  const int64_t n_from_next = (has_next && nlo > 0) ? blocks_len(x, nb - nlo, nlo) : 0;
  char* from_prev = d.halo_tmp;
  char* from_next = d.halo_tmp + (((size_t)n_from_prev * es + 255) & ~(size_t)255);
  if (n_from_prev) add_inplace(x->dtype, x->block_ptr(0), from_prev, n_from_prev, ctx().stream);
  if (n_from_next) add_inplace(x->dtype, x->block_ptr(nb - nlo), from_next, n_from_next, ctx().stream);
}
}  // namespace

int jets_dist_halo_reduce(jets_buf x, int32_t nlo, jets_buf lo, int32_t nhi, jets_buf hi) {
  return guard([&] {
    require_ready(); need_nccl();
    halo_reduce_xfer(x, nlo, lo, nhi, hi);
    halo_reduce_add(x, nlo, nhi);
  });
}
// ---- peer-memory registration (CUDA IPC) ------------------------------------------------------
int jets_dist_register(jets_buf x) {
  return guard([&] {
    require_ready(); need_nccl();
    Dist& d = dist();
    JETS_CHECK(x && x->st && x->st->alloc, JETS_ERR_INVALID, "jets_dist_register needs a library-owned buffer");
    const void* base = alloc_base(x);
    if (d.peers.count(base)) return;
    cudaIpcMemHandle_t mine;
    CUDA_TRY(cudaIpcGetMemHandle(&mine, const_cast<void*>(base)));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    char* dev = nullptr;
    CUDA_TRY(cudaMalloc(&dev, (size_t)(d.size + 1) * 64));
    CUDA_TRY(cudaMemcpyAsync(dev + (size_t)d.size * 64, &mine, 64, cudaMemcpyHostToDevice, ctx().stream));
    NCCL_TRY(d.n.AllGather(dev + (size_t)d.size * 64, dev, 16, ncclFloat32, d.comm, ctx().stream));
    std::vector<cudaIpcMemHandle_t> all(d.size);
    CUDA_TRY(cudaMemcpyAsync(all.data(), dev, (size_t)d.size * 64, cudaMemcpyDeviceToHost, ctx().stream));
    CUDA_TRY(cudaStreamSynchronize(ctx().stream));
    cudaFree(dev);
    Dist::Peer p;
    p.bytes = x->st->bytes;
    if (d.rank > 0) CUDA_TRY(cudaIpcOpenMemHandle(&p.prev, all[d.rank - 1], cudaIpcMemLazyEnablePeerAccess));
    if (d.rank + 1 < d.size) CUDA_TRY(cudaIpcOpenMemHandle(&p.next, all[d.rank + 1], cudaIpcMemLazyEnablePeerAccess));
    d.peers[base] = p;
  });
}

// ---- forward halo gather, split so that it overlaps the interior rows ---------------------------
int jets_dist_halo_exchange_begin(jets_buf x, int32_t nlo, jets_buf lo, int32_t nhi, jets_buf hi) {
  return guard([&] {
    require_ready(); need_nccl();
    Dist& d = dist();
    Context& c = ctx();
    const int nb = x->nblocks();
    const size_t es = dsize(x->dtype);
    const bool has_prev = d.rank > 0, has_next = d.rank + 1 < d.size;
    d.pending_nccl = false;
    if (Dist::Peer* p = peer_of(x)) {
      // pull the previous rank's last nlo own blocks and the next rank's first nhi own blocks with the
      // copy engines over NVLink: no SM is taken from the compute kernel that runs meanwhile
      const size_t n_prev = (has_prev && nlo > 0 && lo) ? (size_t)lo->length() * es : 0;
      const size_t n_next = (has_next && nhi > 0 && hi) ? (size_t)hi->length() * es : 0;
      pull_begin(n_prev ? lo->ptr() : nullptr, n_prev ? remote(p->prev, x, x->block_ptr(nb - nlo)) : nullptr, n_prev,
                 n_next ? hi->ptr() : nullptr, n_next ? remote(p->next, x, x->block_ptr(0)) : nullptr, n_next);
      return;
    }
    // NCCL fallback on a side stream (serialises with kernels of a different shared-memory carve-out)
    ensure_copy_streams();
    CUDA_TRY(cudaEventRecord(d.ev_begin, c.stream));
    CUDA_TRY(cudaStreamWaitEvent(d.copy[0], d.ev_begin, 0));
    cudaStream_t keep = c.stream;
    c.stream = d.copy[0];
    const int rc = jets_dist_halo_exchange(x, nlo, lo, nhi, hi);
    c.stream = keep;
    if (rc != JETS_OK) throw Fail{rc};
    CUDA_TRY(cudaEventRecord(d.ev_copy[0], d.copy[0]));
    d.pending_nccl = true;
  });
}
int jets_dist_halo_exchange_end(void) {
  return guard([&] {
    require_ready(); need_dist();
    Dist& d = dist();
    if (d.pending_nccl) {
      CUDA_TRY(cudaStreamWaitEvent(ctx().stream, d.ev_copy[0], 0));
      d.pending_nccl = false;
      return;
    }
    pull_end();
    stream_barrier();   // nobody may overwrite what a neighbour is still pulling
  });
}

int jets_dist_halo_reduce_begin(jets_buf x, int32_t nlo, jets_buf lo, int32_t nhi, jets_buf hi) {
  return guard([&] {
    require_ready(); need_nccl();
    Dist& d = dist();
    Context& c = ctx();
    d.pending_nccl = false;
    if (Dist::Peer* p = peer_of(x)) {
      const int nb = x->nblocks();
      const size_t es = dsize(x->dtype);
      const bool has_prev = d.rank > 0, has_next = d.rank + 1 < d.size;
      const int64_t n_from_prev = (has_prev && nhi > 0) ? blocks_len(x, 0, nhi) : 0;
      const int64_t n_from_next = (has_next && nlo > 0) ? blocks_len(x, nb - nlo, nlo) : 0;
      const size_t need = (size_t)(n_from_prev + n_from_next) * es + 512;
      if (need > d.halo_tmp_bytes) {
        CUDA_TRY(cudaDeviceSynchronize());
        if (d.halo_tmp) cudaFree(d.halo_tmp);
        CUDA_TRY(cudaMalloc(&d.halo_tmp, need));
        d.halo_tmp_bytes = need;
      }
      char* from_prev = d.halo_tmp;
      char* from_next = d.halo_tmp + (((size_t)n_from_prev * es + 255) & ~(size_t)255);
      JETS_CHECK((!n_from_prev || hi) && (!n_from_next || lo), JETS_ERR_INVALID, "halo views are required on the peer-memory path");
      // the previous rank's partial for my first blocks sits in ITS hi halo (same offset as mine),
      // the next rank's partial for my last blocks in ITS lo halo
      pull_begin(from_prev, n_from_prev ? remote(p->prev, x, hi->ptr()) : nullptr, (size_t)n_from_prev * es,
                 from_next, n_from_next ? remote(p->next, x, lo->ptr()) : nullptr, (size_t)n_from_next * es);
      return;
    }
    ensure_copy_streams();
    CUDA_TRY(cudaEventRecord(d.ev_begin, c.stream));
    CUDA_TRY(cudaStreamWaitEvent(d.copy[0], d.ev_begin, 0));
    cudaStream_t keep = c.stream;
    c.stream = d.copy[0];
    try { halo_reduce_xfer(x, nlo, lo, nhi, hi); } catch (...) { c.stream = keep; throw; }
    c.stream = keep;
    CUDA_TRY(cudaEventRecord(d.ev_copy[0], d.copy[0]));
    d.pending_nccl = true;
  });
}
int jets_dist_halo_reduce_end(jets_buf x, int32_t nlo, int32_t nhi) {
  return guard([&] {
    require_ready(); need_dist();
    Dist& d = dist();
    if (d.pending_nccl) {
      CUDA_TRY(cudaStreamWaitEvent(ctx().stream, d.ev_copy[0], 0));
      d.pending_nccl = false;
    } else {
      pull_end();
      stream_barrier();
    }
    halo_reduce_add(x, nlo, nhi);
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
