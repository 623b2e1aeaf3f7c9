// Distributed operators behind ONE call per apply (include/jets_b200.h: jets_dist_op_*).
//
// Block-banded JopBlock, block rows partitioned over the ranks: rank g holds its rows as
// A_loc, an nloc x (nloc + 2h) JopBlock over [h blocks of rank g-1 | nloc own blocks | h blocks of rank g+1],
// and the matching shards of the domain and range vectors (nloc blocks each).  The reference computes a
// block row as d_r = sum_c A_rc m_c (src/Jets.jl:1015-1030) and a block column of the adjoint as
// m_c = sum_r A_rc' d_r (:1039-1055); both need h blocks of the neighbouring ranks.
//
// Every apply is ONE launch of the fused bundle kernel per rank -- no NCCL call, no copy-engine stream, no
// host synchronisation.  The neighbours' data moves through an "exchange arena" (one cudaMalloc per rank,
// mapped into the two neighbouring processes with CUDA IPC):
//
//   forward   the launch's FIRST units copy this rank's first / last h own blocks into the neighbours' halo
//             buffers with plain 128-bit stores to peer memory (NVLink) and raise the neighbour's "ready" flag
//             (st.release.sys by the CTA that finishes the last such unit); the units of the rows that READ
//             a halo buffer are enumerated LAST, start once this rank's own "ready" flag has been raised
//             (ld.acquire.sys poll in the producer warp -- normally long satisfied) and report "done" to the
//             writer, which may then reuse the buffer (double buffered: the wait is for epoch e-2).
//   adjoint   the partial sums this rank's rows contribute to the neighbours' columns are stored straight
//             into the neighbours' staging buffers; the owner adds them as one more term of its row sum, the
//             previous rank's partial first and the next rank's last -- the single-GPU order of :1049, so the
//             result of a block-tridiagonal operator is bit-identical to the single-GPU apply.
//
// Flags are epoch counters (one per direction of use), so nothing is ever reset and a late rank simply finds
// its flags already raised.  The same plans, cut into block-row chunks, drive the host-buffer pipeline
// (jets_dist_apply_normal_host): upload k | forward k-1, adjoint k-2 | download k-2 on three streams.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>
#include "dist.hpp"

namespace jets {
extern uint64_t g_epoch;
namespace {

constexpr int kMaxHalo = 4;

struct ArenaLayout {                 // byte offsets from the arena base; exchanged with the neighbours
  int64_t flags;                     // 2 sets (forward, adjoint) x kGateFlags flag words, 128 B apart
  int64_t buf[2][2][2];              // [forward|adjoint][lo|hi][parity]
  int64_t len_first[kMaxHalo], len_last[kMaxHalo];   // my first / last h own block lengths
  int64_t len_lo[kMaxHalo], len_hi[kMaxHalo];        // the halo block lengths A_loc declares
  int64_t nloc, halo, dtype;
};
struct Exchange {
  cudaIpcMemHandle_t handle;
  ArenaLayout lay;
};

void ck(int rc) {
  if (rc != JETS_OK) throw Fail{rc};
}
jets_op strip_op(jets_op a) {
  while (a->kind == K_LNVIEW || a->kind == K_ADJ) a = a->kids[0];
  return a;
}

struct HostPipe;

}  // namespace
}  // namespace jets

using namespace jets;

struct jets_dist_op_s {
  int kind = 0;                      // 0 banded (peer-memory halos), 1 dense structure (all-gather / reduce-scatter)
  jets_op A = nullptr;               // retained
  int halo = 0, nloc = 0, dtype = JETS_F32;
  bool has_prev = false, has_next = false;
  char* arena = nullptr;
  size_t arena_bytes = 0;
  ArenaLayout lay{}, prev_lay{}, next_lay{};
  char* prev_base = nullptr;         // the neighbours' arenas mapped into this process
  char* next_base = nullptr;
  bool loopback = false;             // single-process test mode: prev_base == next_base == arena
  uint32_t epoch[2] = {0, 0};        // forward / adjoint applies issued so far
  std::shared_ptr<Plan> mono[2];
  std::shared_ptr<Plan> mono_pull;   // forward on a registered input vector: halo terms read the neighbours' memory
  std::unique_ptr<HostPipe> pipe;
  // registered domain vectors (jets_dist_op_register): where the neighbours' copies are mapped
  struct Reg { char* prev = nullptr; char* next = nullptr; void* prev_map = nullptr; void* next_map = nullptr; jets_buf keep = nullptr; };
  std::map<const void*, Reg> regs;   // keyed by the local vector's first byte
  int64_t n_own = 0, n_rng = 0;      // elements of the own domain / range shards
  // dense structure
  jets_buf full = nullptr;           // the whole domain (all-gathered forward, partial sums adjoint)
  int64_t shard_len = 0;
};

namespace jets {
namespace {

std::set<jets_dist_op>& live_ops() {
  static std::set<jets_dist_op> s;
  return s;
}

uint32_t* my_flags_base(jets_dist_op D) { return reinterpret_cast<uint32_t*>(D->arena + D->lay.flags); }
uint64_t gate_timeout_ns() {   // JETS_B200_GATE_TIMEOUT_MS: how long a work unit waits for a neighbour's flag (default 30 s; 0 = forever)
  static const uint64_t v = [] {
    const char* e = getenv("JETS_B200_GATE_TIMEOUT_MS");
    return (uint64_t)(e ? atoll(e) : 30000) * 1000000ull;
  }();
  return v;
}

GateLaunch gate_for(jets_dist_op D, int adj, uint32_t e, const jets_dist_op_s::Reg* pull = nullptr) {
  GateLaunch g;
  if (!D->has_prev && !D->has_next) return g;
  const int par = e & 1;
  uint32_t* my_flags = my_flags_base(D) + adj * kGateFlags * kGateFlagStride;
  g.flags = my_flags;
  g.wait_val[GF_LO_READY] = e;
  g.wait_val[GF_HI_READY] = e;
  g.wait_val[GF_PREV_DONE] = e - 2;    // double buffered: the buffer written now was last read at epoch e-2
  g.wait_val[GF_NEXT_DONE] = e - 2;
  for (int k = 0; k < kGateFlags; ++k) g.sig_val[k] = e;
  g.err = my_flags_base(D) + 2 * kGateFlags * kGateFlagStride;
  g.timeout_ns = gate_timeout_ns();
  g.in_alt[0] = D->arena + D->lay.buf[adj][0][par];
  g.in_alt[1] = D->arena + D->lay.buf[adj][1][par];
  if (D->has_prev) {
    uint32_t* pf = reinterpret_cast<uint32_t*>(D->prev_base + D->prev_lay.flags) + adj * kGateFlags * kGateFlagStride;
    g.sig_addr[GS_PREV_HI_READY] = pf + GF_HI_READY * kGateFlagStride;
    g.sig_addr[GS_PREV_NEXT_DONE] = pf + GF_NEXT_DONE * kGateFlagStride;
    g.out_alt[0] = D->prev_base + D->prev_lay.buf[adj][1][par];      // the previous rank's hi buffer
  }
  if (D->has_next) {
    uint32_t* nf = reinterpret_cast<uint32_t*>(D->next_base + D->next_lay.flags) + adj * kGateFlags * kGateFlagStride;
    g.sig_addr[GS_NEXT_LO_READY] = nf + GF_LO_READY * kGateFlagStride;
    g.sig_addr[GS_NEXT_PREV_DONE] = nf + GF_PREV_DONE * kGateFlagStride;
    g.out_alt[1] = D->next_base + D->next_lay.buf[adj][0][par];      // the next rank's lo buffer
  }
  if (pull) {
    // the halo terms read the neighbours' vectors in place; the launch may not end before the neighbours have
    // finished reading this rank's vector
    g.in_alt[0] = pull->prev;
    g.in_alt[1] = pull->next;
    g.exit_wait = (D->has_prev ? 1 << GF_PREV_DONE : 0) | (D->has_next ? 1 << GF_NEXT_DONE : 0);
    g.exit_val[GF_PREV_DONE] = g.exit_val[GF_NEXT_DONE] = e;
  }
  return g;
}

int owned_signals(jets_dist_op D) {
  return (D->has_prev ? (1 << GS_PREV_HI_READY) | (1 << GS_PREV_NEXT_DONE) : 0) |
         (D->has_next ? (1 << GS_NEXT_LO_READY) | (1 << GS_NEXT_PREV_DONE) : 0);
}

bool buf_ok(jets_buf x) { return x->guarded() && (reinterpret_cast<uintptr_t>(x->ptr()) & 15) == 0; }

void launch_plan(jets_dist_op D, Plan& p, const char* in, char* out, const GateLaunch& g, cudaStream_t s) {
  for (Step& st : p.steps) {
    JETS_CHECK(st.kind == ST_FUSED && st.fused.bundle, JETS_ERR_INVALID, "internal: banded plan with a non-bundle step");
    launch_fused_bundle(st.fused, D->dtype, in, out, s, nullptr, st.fused.gated ? &g : nullptr);
  }
}

// ------------------------------------------------------------------ host-buffer pipeline ------
// m_host -> d = A m -> m' = A' d -> m'_host over block-row chunks on three streams: the upload of chunk k
// overlaps the forward of chunk k-1 and the adjoint of chunk k-2, whose result goes back to the host on the
// third stream.  Each chunk is a banded plan over a row range of the SAME operator (so every output element
// is computed by the same kernel arithmetic as in the monolithic apply: bit-identical).  With neighbours,
// the chunks that touch a halo are issued last: their kernels wait in place for the neighbour's flag.
struct HostPipe {
  int K = 0;
  std::vector<std::pair<int, int>> chunks;       // [a, b) block rows / own columns
  std::vector<std::shared_ptr<Plan>> fwd, adj;
  std::shared_ptr<Plan> push[2], part[2];        // forward pushes / adjoint partial sums towards prev, next
  std::vector<int> up_need;                      // forward chunk k needs uploads 0..up_need[k]
  std::vector<std::vector<int>> fw_need;         // adjoint chunk j needs these forward chunks
  struct Item { int what, k; };                  // 0 fwd, 1 adj, 2 push_prev, 3 push_next, 4 part_prev, 5 part_next
  std::vector<Item> seq;
  std::vector<int> down_order;
  jets_buf x = nullptr, d = nullptr, m = nullptr;
  cudaStream_t s_up = nullptr, s_comp = nullptr, s_down = nullptr;
  std::vector<cudaEvent_t> ev_up, ev_adj;
  cudaEvent_t ev_call = nullptr, ev_comp = nullptr, ev_down = nullptr;
  bool have_prev_step = false;
  ~HostPipe() {
    for (cudaStream_t s : {s_up, s_comp, s_down})
      if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (auto e : ev_up) cudaEventDestroy(e);
    for (auto e : ev_adj) cudaEventDestroy(e);
    for (auto e : {ev_call, ev_comp, ev_down})
      if (e) cudaEventDestroy(e);
    for (jets_buf b : {x, d, m})
      if (b) jets_buf_destroy(b);
  }
};

bool block_nonzero(jets_op A, int r, int j) { return strip_op(A->kids[r + (size_t)j * A->R])->kind != K_ZERO; }

void build_pipe_plans(jets_dist_op D, HostPipe& P) {
  const int h = D->halo, n = D->nloc;
  BandedSel base;
  base.has_prev = D->has_prev; base.has_next = D->has_next;
  P.fwd.clear(); P.adj.clear();
  for (auto& c : P.chunks) {
    BandedSel s = base;
    s.mode = JETS_MODE_DF; s.row_begin = c.first; s.row_end = c.second;
    // the chunk that holds the rows reading a halo owns the matching "done" signal
    if (D->has_prev && c.first == 0) s.owned |= 1 << GS_PREV_NEXT_DONE;
    if (D->has_next && c.second == n) s.owned |= 1 << GS_NEXT_PREV_DONE;
    P.fwd.push_back(build_banded_plan(D->A, h, s));
    s.mode = JETS_MODE_DFT;
    P.adj.push_back(build_banded_plan(D->A, h, s));
  }
  for (int side = 0; side < 2; ++side) {
    P.push[side].reset(); P.part[side].reset();
    if (!(side == 0 ? D->has_prev : D->has_next)) continue;
    BandedSel s = base;
    s.row_begin = s.row_end = 0;
    s.send_prev = side == 0; s.send_next = side == 1;
    s.owned = 1 << (side == 0 ? GS_PREV_HI_READY : GS_NEXT_LO_READY);
    s.mode = JETS_MODE_DF;
    P.push[side] = build_banded_plan(D->A, h, s);
    s.mode = JETS_MODE_DFT;
    P.part[side] = build_banded_plan(D->A, h, s);
  }
}

// Chunk bounds, dependencies and the issue order of the compute stream from the block structure alone
// (nz(r, j): block (r, j) of the nloc x (nloc + 2h) rank-local operator is not a zero block).  Pure host code:
// exported as jets_dist_pipeline_schedule so that the CPU tests can replay it.
template <class NZ>
void pipe_schedule(HostPipe& P, int n, int h, int nchunks, bool has_prev, bool has_next, NZ nz) {
  const int hh = std::max(1, h);
  nchunks = std::max(1, std::min(nchunks, n / hh));
  for (int k = 0; k < nchunks; ++k) {
    const int a = (int)((int64_t)n * k / nchunks), b = (int)((int64_t)n * (k + 1) / nchunks);
    if (b > a) P.chunks.push_back({a, b});
  }
  const int K = P.K = (int)P.chunks.size();
  auto chunk_of = [&](int blk) {
    for (int k = 0; k < K; ++k)
      if (blk >= P.chunks[k].first && blk < P.chunks[k].second) return k;
    return K - 1;
  };
  // forward chunk k reads own blocks -> uploads; adjoint chunk j reads the range blocks of the rows that hit its
  // columns -> forward chunks.  "Late" = touches a neighbour's data (its kernel waits for a flag).
  P.up_need.assign(K, 0);
  P.fw_need.assign(K, {});
  std::vector<char> late_f(K, 0), late_a(K, 0);
  for (int k = 0; k < K; ++k) {
    int need = k;
    for (int r = P.chunks[k].first; r < P.chunks[k].second; ++r)
      for (int j = 0; j < n + 2 * h; ++j) {
        if (!nz(r, j)) continue;
        if (j < h) { if (has_prev) late_f[k] = 1; }
        else if (j >= n + h) { if (has_next) late_f[k] = 1; }
        else need = std::max(need, chunk_of(j - h));
      }
    P.up_need[k] = need;
  }
  for (int jn = 0; jn < K; ++jn) {
    std::set<int> need;
    need.insert(jn);
    for (int b = P.chunks[jn].first; b < P.chunks[jn].second; ++b)
      for (int r = 0; r < n; ++r)
        if (nz(r, b + h)) need.insert(chunk_of(r));
    P.fw_need[jn].assign(need.begin(), need.end());
    for (int f : need) late_a[jn] |= late_f[f];
    if (has_prev && P.chunks[jn].first < h) late_a[jn] = 1;
    if (has_next && P.chunks[jn].second > n - h) late_a[jn] = 1;
  }
  std::vector<char> f_done(K, 0), a_done(K, 0);
  auto try_adj = [&]() {
    for (int jn = 0; jn < K; ++jn) {
      if (a_done[jn] || late_a[jn]) continue;
      bool ok = true;
      for (int f : P.fw_need[jn]) ok = ok && f_done[f];
      if (ok) { P.seq.push_back({1, jn}); a_done[jn] = 1; P.down_order.push_back(jn); }
    }
  };
  // my first h blocks go to the previous rank as soon as they are uploaded; everything that does not touch a
  // neighbour streams k / k-1 / k-2; the last h blocks can only be pushed once the upload is complete, and the
  // chunks that wait for a neighbour follow, forward before the partial sums before the adjoint
  if (has_prev) P.seq.push_back({2, 0});
  for (int k = 0; k < K; ++k) {
    if (late_f[k]) continue;
    P.seq.push_back({0, k});
    f_done[k] = 1;
    try_adj();
  }
  if (has_next) P.seq.push_back({3, 0});
  for (int k = 0; k < K; ++k)
    if (late_f[k]) { P.seq.push_back({0, k}); f_done[k] = 1; }
  if (has_prev) P.seq.push_back({4, 0});
  if (has_next) P.seq.push_back({5, 0});
  for (int jn = 0; jn < K; ++jn)
    if (!a_done[jn]) { P.seq.push_back({1, jn}); a_done[jn] = 1; P.down_order.push_back(jn); }
}

bool pipe_plans_valid(const HostPipe& P) {
  for (auto* v : {&P.fwd, &P.adj})
    for (auto& p : *v)
      if (!p->valid()) return false;
  for (auto& p : {P.push[0], P.push[1], P.part[0], P.part[1]})
    if (p && !p->valid()) return false;
  return true;
}

void make_pipe(jets_dist_op D, int nchunks) {
  const int h = D->halo, n = D->nloc;
  auto P = std::make_unique<HostPipe>();
  pipe_schedule(*P, n, h, nchunks, D->has_prev, D->has_next, [&](int r, int j) { return block_nonzero(D->A, r, j); });
  const int K = P->K;
  build_pipe_plans(D, *P);
  // work vectors (own shards) and streams
  std::vector<int64_t> own(n), rng(n);
  for (int b = 0; b < n; ++b) { own[b] = D->A->dom.len[h + b]; rng[b] = D->A->rng.len[b]; }
  ck(jets_buf_create((jets_dtype)D->dtype, n, own.data(), &P->x));
  ck(jets_buf_create((jets_dtype)D->dtype, n, rng.data(), &P->d));
  ck(jets_buf_create((jets_dtype)D->dtype, n, own.data(), &P->m));
  CUDA_TRY(cudaStreamSynchronize(ctx().stream));   // the zero fills of the work vectors
  for (cudaStream_t* s : {&P->s_up, &P->s_comp, &P->s_down}) CUDA_TRY(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
  P->ev_up.resize(K); P->ev_adj.resize(K);
  for (int k = 0; k < K; ++k) {
    CUDA_TRY(cudaEventCreateWithFlags(&P->ev_up[k], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&P->ev_adj[k], cudaEventDisableTiming));
  }
  for (cudaEvent_t* e : {&P->ev_call, &P->ev_comp, &P->ev_down}) CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  D->pipe = std::move(P);
}

void pipe_step(jets_dist_op D, char* host_out, const char* host_in) {
  HostPipe& P = *D->pipe;
  if (!pipe_plans_valid(P)) build_pipe_plans(D, P);
  const size_t esz = dsize(D->dtype);
  const int K = P.K;
  // ordered after whatever the caller issued on the context stream
  CUDA_TRY(cudaEventRecord(P.ev_call, ctx().stream));
  for (cudaStream_t s : {P.s_up, P.s_comp, P.s_down}) CUDA_TRY(cudaStreamWaitEvent(s, P.ev_call, 0));
  if (P.have_prev_step) {
    CUDA_TRY(cudaStreamWaitEvent(P.s_up, P.ev_comp, 0));     // x may be overwritten once the previous step's compute is done
    CUDA_TRY(cudaStreamWaitEvent(P.s_comp, P.ev_down, 0));   // m once the previous step's downloads are done
  }
  for (int k = 0; k < K; ++k) {
    const int64_t o = P.x->blk_off[P.chunks[k].first], e = P.x->blk_off[P.chunks[k].second];
    CUDA_TRY(cudaMemcpyAsync(P.x->ptr() + o * esz, host_in + o * esz, (size_t)(e - o) * esz, cudaMemcpyHostToDevice, P.s_up));
    CUDA_TRY(cudaEventRecord(P.ev_up[k], P.s_up));
  }
  const uint32_t ef = ++D->epoch[0], ea = ++D->epoch[1];
  const GateLaunch gf = gate_for(D, 0, ef), ga = gate_for(D, 1, ea);
  const int h = D->halo, n = D->nloc;
  auto up_chunk_of_block = [&](int blk) {
    for (int k = 0; k < K; ++k)
      if (blk < P.chunks[k].second) return k;
    return K - 1;
  };
  for (const HostPipe::Item& it : P.seq) {
    switch (it.what) {
      case 0:
        CUDA_TRY(cudaStreamWaitEvent(P.s_comp, P.ev_up[P.up_need[it.k]], 0));
        launch_plan(D, *P.fwd[it.k], P.x->ptr(), P.d->ptr(), gf, P.s_comp);
        break;
      case 1:
        launch_plan(D, *P.adj[it.k], P.d->ptr(), P.m->ptr(), ga, P.s_comp);
        CUDA_TRY(cudaEventRecord(P.ev_adj[it.k], P.s_comp));
        break;
      case 2:   // my first h blocks -> the previous rank's hi halo, as soon as they are uploaded
        CUDA_TRY(cudaStreamWaitEvent(P.s_comp, P.ev_up[up_chunk_of_block(h - 1)], 0));
        launch_plan(D, *P.push[0], P.x->ptr(), P.d->ptr(), gf, P.s_comp);
        break;
      case 3:
        CUDA_TRY(cudaStreamWaitEvent(P.s_comp, P.ev_up[up_chunk_of_block(n - 1)], 0));
        launch_plan(D, *P.push[1], P.x->ptr(), P.d->ptr(), gf, P.s_comp);
        break;
      case 4: launch_plan(D, *P.part[0], P.d->ptr(), P.m->ptr(), ga, P.s_comp); break;
      case 5: launch_plan(D, *P.part[1], P.d->ptr(), P.m->ptr(), ga, P.s_comp); break;
    }
  }
  CUDA_TRY(cudaEventRecord(P.ev_comp, P.s_comp));
  for (int jn : P.down_order) {
    const int64_t o = P.m->blk_off[P.chunks[jn].first], e = P.m->blk_off[P.chunks[jn].second];
    CUDA_TRY(cudaStreamWaitEvent(P.s_down, P.ev_adj[jn], 0));
    CUDA_TRY(cudaMemcpyAsync(host_out + o * esz, P.m->ptr() + o * esz, (size_t)(e - o) * esz, cudaMemcpyDeviceToHost, P.s_down));
  }
  CUDA_TRY(cudaEventRecord(P.ev_down, P.s_down));
  P.have_prev_step = true;
}

// ------------------------------------------------------------------ dense block structure ------
// Every block row needs the WHOLE domain (src/Jets.jl:1015-1030 with no zero blocks) and every rank's rows
// contribute to every block column of the adjoint (:1039-1055): the forward all-gathers the domain shards,
// the adjoint reduce-scatters the per-rank partial sums (NCCL over NVLink; the shards are equal-sized).
// complex shards travel as twice as many reals (a sum of complex numbers is the sum of the parts)
int nccl_dtype(int dt) { return real_of(dt) == JETS_F32 ? ncclFloat32 : ncclFloat64; }
size_t nccl_count(int dt, int64_t n) { return (size_t)n * (is_cplx(dt) ? 2 : 1); }

void dense_apply(jets_dist_op D, int mode, jets_buf out, jets_buf in) {
  need_nccl();
  Dist& d = dist();
  const int adj = mode == JETS_MODE_DFT;
  const int64_t nrng = D->A->rng.total();
  if (!adj) {
    JETS_CHECK(in->length() == D->shard_len, JETS_ERR_SHAPE, "input shard has %lld elements, expected %lld (domain / ranks)", (long long)in->length(), (long long)D->shard_len);
    JETS_CHECK(out->length() == nrng, JETS_ERR_SHAPE, "output has %lld elements, the rank-local rows produce %lld", (long long)out->length(), (long long)nrng);
    NCCL_TRY(d.n.AllGather(in->ptr(), D->full->ptr(), nccl_count(D->dtype, D->shard_len), nccl_dtype(D->dtype), d.comm, ctx().stream));
    ck(jets_apply(D->A, mode == JETS_MODE_F && !D->A->linear ? JETS_MODE_F : JETS_MODE_DF, out, D->full, 0));
  } else {
    JETS_CHECK(in->length() == nrng, JETS_ERR_SHAPE, "input has %lld elements, the rank-local rows take %lld", (long long)in->length(), (long long)nrng);
    JETS_CHECK(out->length() == D->shard_len, JETS_ERR_SHAPE, "output shard has %lld elements, expected %lld (domain / ranks)", (long long)out->length(), (long long)D->shard_len);
    ck(jets_apply(D->A, JETS_MODE_DFT, D->full, in, 0));
    NCCL_TRY(d.n.ReduceScatter(D->full->ptr(), out->ptr(), nccl_count(D->dtype, D->shard_len), nccl_dtype(D->dtype), ncclSum, d.comm, ctx().stream));
  }
}

void close_regs(jets_dist_op D) {
  for (auto& kv : D->regs) {
    if (!D->loopback) {
      if (kv.second.prev_map) cudaIpcCloseMemHandle(kv.second.prev_map);
      if (kv.second.next_map) cudaIpcCloseMemHandle(kv.second.next_map);
    }
    if (kv.second.keep) jets_buf_destroy(kv.second.keep);   // the registration held the vector alive (its address is the key)
  }
  D->regs.clear();
}

void destroy_op(jets_dist_op D) {
  if (D->pipe) D->pipe.reset();
  cudaStreamSynchronize(ctx().stream);
  D->mono[0].reset(); D->mono[1].reset();
  D->mono_pull.reset();
  close_regs(D);
  if (D->prev_base && !D->loopback) cudaIpcCloseMemHandle(D->prev_base);
  if (D->next_base && !D->loopback) cudaIpcCloseMemHandle(D->next_base);
  if (D->arena) cudaFree(D->arena);
  if (D->full) jets_buf_destroy(D->full);
  if (D->A) jets_op_destroy(D->A);
  live_ops().erase(D);
  delete D;
}

}  // namespace

void dist_ops_shutdown() {
  // NCCL is going away: release the peer mappings while the neighbours are still alive
  for (jets_dist_op D : live_ops()) {
    if (D->pipe) D->pipe.reset();
    if (D->loopback) continue;
    D->mono_pull.reset();
    close_regs(D);
    if (D->prev_base) { cudaIpcCloseMemHandle(D->prev_base); D->prev_base = nullptr; }
    if (D->next_base) { cudaIpcCloseMemHandle(D->next_base); D->next_base = nullptr; }
    D->has_prev = D->has_next = false;
    D->mono[0].reset(); D->mono[1].reset();
  }
}

}  // namespace jets

extern "C" {

int jets_dist_op_create(jets_op A_loc, int32_t halo, jets_dist_op* out) {
  return guard([&] {
    require_ready();
    JETS_CHECK(A_loc && A_loc->refs > 0 && out, JETS_ERR_INVALID, "null or destroyed operator handle");
    JETS_CHECK(halo >= 0 && halo <= kMaxHalo, JETS_ERR_INVALID, "halo width must be 0..%d", kMaxHalo);
    JETS_CHECK(A_loc->kind == K_BLOCK, JETS_ERR_INVALID, "jets_dist_op_create expects the rank-local block rows as a JopBlock");
    JETS_CHECK(A_loc->C == A_loc->R + 2 * halo, JETS_ERR_SHAPE,
               "the rank-local operator must be nloc x (nloc + 2*halo) blocks over the halo-extended domain, got %d x %d with halo %d",
               A_loc->R, A_loc->C, halo);
    JETS_CHECK(A_loc->R >= std::max(1, (int)halo), JETS_ERR_SHAPE, "fewer local block rows than the halo width");
    JETS_CHECK(!is_cplx(A_loc->dtype), JETS_ERR_UNSUPPORTED, "distributed banded apply: complex eltypes are not implemented");
    Dist& d = dist();
    std::unique_ptr<jets_dist_op_s> D(new jets_dist_op_s());
    D->halo = halo; D->nloc = A_loc->R; D->dtype = A_loc->dtype;
    D->has_prev = d.ready && d.rank > 0 && halo > 0;
    D->has_next = d.ready && d.rank + 1 < d.size && halo > 0;
    // JETS_B200_DIST_LOOPBACK=1 (single process): the rank is its own previous and next neighbour -- a block-
    // CIRCULANT operator -- so that the whole gated path (push units, flag waits, signals) runs, and can be
    // profiled, on one GPU without a second process
    const bool loopback = !(d.ready && d.size > 1) && halo > 0 && getenv("JETS_B200_DIST_LOOPBACK") && atoi(getenv("JETS_B200_DIST_LOOPBACK"));
    if (loopback) D->has_prev = D->has_next = true;
    const int n = D->nloc, h = halo;
    const size_t esz = dsize(D->dtype);
    for (int b = 0; b < n; ++b) { D->n_own += A_loc->dom.len[h + b]; D->n_rng += A_loc->rng.len[b]; }
    // arena: flag words, then the halo (forward) and staging (adjoint) buffers, each with readable guards
    ArenaLayout& L = D->lay;
    L.nloc = n; L.halo = h; L.dtype = D->dtype;
    int64_t lo = 0, hi = 0, first = 0, last = 0;
    for (int k = 0; k < h; ++k) {
      L.len_lo[k] = A_loc->dom.len[k]; L.len_hi[k] = A_loc->dom.len[n + h + k];
      L.len_first[k] = A_loc->dom.len[h + k]; L.len_last[k] = A_loc->dom.len[n + k];
      lo += L.len_lo[k]; hi += L.len_hi[k]; first += L.len_first[k]; last += L.len_last[k];
    }
    size_t off = 0;
    L.flags = 0;
    off += (2 * kGateFlags + 1) * kGateFlagStride * sizeof(uint32_t);   // + the wait-timeout counter
    auto place = [&](int64_t elems) {
      off += kGuardBytes;
      const int64_t o = (int64_t)off;
      off += (((size_t)elems * esz + 255) & ~(size_t)255) + kGuardBytes;
      return o;
    };
    for (int par = 0; par < 2; ++par) {
      L.buf[0][0][par] = place(lo);      // forward: the previous rank's last h blocks
      L.buf[0][1][par] = place(hi);      //          the next rank's first h blocks
      L.buf[1][0][par] = place(first);   // adjoint: the previous rank's partial sums for my first h columns
      L.buf[1][1][par] = place(last);    //          the next rank's for my last h columns
    }
    if (D->has_prev || D->has_next) {
      D->arena_bytes = off;
      CUDA_TRY(cudaMalloc(&D->arena, off));
      CUDA_TRY(cudaMemset(D->arena, 0, off));
      CUDA_TRY(cudaDeviceSynchronize());
    }
    if (d.ready && d.size > 1) {
      // collective: every rank publishes its arena (IPC handle + layout); neighbours map each other
      Exchange mine{};
      if (D->arena) CUDA_TRY(cudaIpcGetMemHandle(&mine.handle, D->arena));
      mine.lay = L;
      std::vector<Exchange> all(d.size);
      dist_allgather_host(&mine, all.data(), sizeof(Exchange));
      auto open = [&](int r, char** base, ArenaLayout* lay) {
        *lay = all[r].lay;
        JETS_CHECK(lay->halo == h && lay->dtype == D->dtype, JETS_ERR_SHAPE, "rank %d built its distributed operator with another halo width or eltype", r);
        void* p = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess));
        *base = reinterpret_cast<char*>(p);
      };
      if (D->has_prev) {
        open(d.rank - 1, &D->prev_base, &D->prev_lay);
        for (int k = 0; k < h; ++k)
          JETS_CHECK(D->prev_lay.len_last[k] == L.len_lo[k] && D->prev_lay.len_hi[k] == L.len_first[k], JETS_ERR_SHAPE,
                     "halo block %d: this rank and rank %d disagree about the block lengths at their common boundary", k, d.rank - 1);
      }
      if (D->has_next) {
        open(d.rank + 1, &D->next_base, &D->next_lay);
        for (int k = 0; k < h; ++k)
          JETS_CHECK(D->next_lay.len_first[k] == L.len_hi[k] && D->next_lay.len_lo[k] == L.len_last[k], JETS_ERR_SHAPE,
                     "halo block %d: this rank and rank %d disagree about the block lengths at their common boundary", k, d.rank + 1);
      }
    }
    if (loopback) {
      D->loopback = true;
      D->prev_base = D->next_base = D->arena;
      D->prev_lay = D->next_lay = L;
    }
    A_loc->refs++;
    D->A = A_loc;
    live_ops().insert(D.get());
    *out = D.release();
  });
}

int jets_dist_op_create_dense(jets_op A_loc, jets_dist_op* out) {
  return guard([&] {
    require_ready(); need_nccl();
    JETS_CHECK(A_loc && A_loc->refs > 0 && out, JETS_ERR_INVALID, "null or destroyed operator handle");
    Dist& d = dist();
    const int64_t total = A_loc->dom.total();
    JETS_CHECK(total % d.size == 0, JETS_ERR_SHAPE, "the domain (%lld elements) does not split into %d equal shards", (long long)total, d.size);
    std::unique_ptr<jets_dist_op_s> D(new jets_dist_op_s());
    D->kind = 1;
    D->dtype = A_loc->dtype;
    D->nloc = (int)A_loc->rng.len.size();
    D->shard_len = total / d.size;
    D->n_own = D->shard_len;
    D->n_rng = A_loc->rng.total();
    ck(jets_buf_create((jets_dtype)D->dtype, (int32_t)A_loc->dom.len.size(), A_loc->dom.len.data(), &D->full));
    A_loc->refs++;
    D->A = A_loc;
    live_ops().insert(D.get());
    *out = D.release();
  });
}

int jets_dist_op_destroy(jets_dist_op D) {
  return guard([&] {
    JETS_CHECK(D && live_ops().count(D), JETS_ERR_INVALID, "null or destroyed distributed operator handle");
    destroy_op(D);
  });
}

int jets_dist_apply(jets_dist_op D, int mode, jets_buf out, jets_buf in) {
  return guard([&] {
    require_ready();
    JETS_CHECK(D && live_ops().count(D), JETS_ERR_INVALID, "null or destroyed distributed operator handle");
    JETS_CHECK(out && out->refs > 0 && in && in->refs > 0, JETS_ERR_INVALID, "null or destroyed buffer handle");
    JETS_CHECK(mode >= 0 && mode <= 2, JETS_ERR_INVALID, "bad mode %d", mode);
    JETS_CHECK(in->dtype == D->dtype && out->dtype == D->dtype, JETS_ERR_DTYPE, "operator eltype %d but in/out eltypes %d/%d", D->dtype,
               in->dtype, out->dtype);
    const int adj = mode == JETS_MODE_DFT;
    if (adj) JETS_CHECK(D->A->linear, JETS_ERR_NOT_LINEAR, "mul!(m, A', d) requires a linear operator (src/Jets.jl:392)");
    if (D->kind == 1) { dense_apply(D, mode, out, in); return; }
    const int64_t nin = adj ? D->n_rng : D->n_own, nout = adj ? D->n_own : D->n_rng;
    JETS_CHECK(in->length() == nin, JETS_ERR_SHAPE, "input shard has %lld elements, the rank-local operator expects %lld", (long long)in->length(), (long long)nin);
    JETS_CHECK(out->length() == nout, JETS_ERR_SHAPE, "output shard has %lld elements, the rank-local operator produces %lld", (long long)out->length(), (long long)nout);
    JETS_CHECK(buf_ok(in) && buf_ok(out), JETS_ERR_UNSUPPORTED, "distributed banded apply: in/out must be library-owned (guarded, 16-byte aligned) vectors");
    if (D->pipe && D->pipe->have_prev_step) {
      // pipelined host-buffer steps of this operator may still be in flight on the library's streams: the flag
      // epochs are shared, so the apply is ordered after them
      CUDA_TRY(cudaStreamWaitEvent(ctx().stream, D->pipe->ev_down, 0));
      CUDA_TRY(cudaStreamWaitEvent(ctx().stream, D->pipe->ev_comp, 0));
    }
    const jets_dist_op_s::Reg* reg = nullptr;
    if (!adj && (D->has_prev || D->has_next)) {
      auto it = D->regs.find(in->ptr());
      if (it != D->regs.end()) reg = &it->second;
    }
    std::shared_ptr<Plan>& plan = reg ? D->mono_pull : D->mono[adj];
    const int pmode = adj ? JETS_MODE_DFT : (mode == JETS_MODE_F && !D->A->linear ? JETS_MODE_F : JETS_MODE_DF);
    if (!plan || !plan->valid()) {
      BandedSel s;
      s.mode = pmode;
      s.row_begin = 0; s.row_end = D->nloc;
      s.send_prev = s.send_next = true;
      s.has_prev = D->has_prev; s.has_next = D->has_next;
      s.owned = owned_signals(D);
      s.pull = reg != nullptr;
      plan = build_banded_plan(D->A, D->halo, s);
    }
    const uint32_t e = ++D->epoch[adj];
    launch_plan(D, *plan, in->ptr(), out->ptr(), gate_for(D, adj, e, reg), ctx().stream);
  });
}

int jets_dist_op_register(jets_dist_op D, jets_buf x) {
  return guard([&] {
    require_ready();
    JETS_CHECK(D && live_ops().count(D), JETS_ERR_INVALID, "null or destroyed distributed operator handle");
    JETS_CHECK(x && x->refs > 0, JETS_ERR_INVALID, "null or destroyed buffer handle");
    JETS_CHECK(D->kind == 0, JETS_ERR_UNSUPPORTED, "jets_dist_op_register applies to block-banded operators");
    JETS_CHECK(x->dtype == D->dtype && x->length() == D->n_own, JETS_ERR_SHAPE, "the vector is not a domain shard of this operator");
    JETS_CHECK(x->st && x->st->alloc && buf_ok(x), JETS_ERR_UNSUPPORTED, "only library-owned vectors can be registered");
    if (D->regs.count(x->ptr())) return;
    Dist& d = dist();
    const size_t esz = dsize(D->dtype);
    const int h = D->halo;
    int64_t lo_elems = 0, first_elems = 0;
    for (int k = 0; k < h; ++k) { lo_elems += D->lay.len_lo[k]; first_elems += D->lay.len_first[k]; }
    jets_dist_op_s::Reg r;
    if (D->loopback) {
      // own previous / next neighbour: the last h blocks sit at the end of the shard, the first h at its start
      r.prev = x->ptr() + (size_t)(D->n_own - lo_elems) * esz;
      r.next = x->ptr();
      x->refs++;
      r.keep = x;
      D->regs[x->ptr()] = r;
      return;
    }
    if (!(d.ready && d.size > 1)) return;     // one rank: nothing to map
    struct Rec { cudaIpcMemHandle_t handle; int64_t data_off, elems; } mine{}, zero{};
    (void)zero;
    CUDA_TRY(cudaIpcGetMemHandle(&mine.handle, x->st->alloc));
    mine.data_off = x->ptr() - reinterpret_cast<char*>(x->st->alloc);
    mine.elems = x->length();
    std::vector<Rec> all(d.size);
    dist_allgather_host(&mine, all.data(), sizeof(Rec));
    auto open = [&](int rk, void** map) {
      CUDA_TRY(cudaIpcOpenMemHandle(map, all[rk].handle, cudaIpcMemLazyEnablePeerAccess));
      return reinterpret_cast<char*>(*map) + all[rk].data_off;
    };
    if (D->has_prev) r.prev = open(d.rank - 1, &r.prev_map) + (size_t)(all[d.rank - 1].elems - lo_elems) * esz;   // its last h blocks
    if (D->has_next) r.next = open(d.rank + 1, &r.next_map);                                                     // its first h blocks
    x->refs++;
    r.keep = x;
    D->regs[x->ptr()] = r;
  });
}

int jets_dist_apply_normal_host(jets_dist_op D, void* host_out, const void* host_in, int32_t nchunks) {
  return guard([&] {
    require_ready();
    JETS_CHECK(D && live_ops().count(D), JETS_ERR_INVALID, "null or destroyed distributed operator handle");
    JETS_CHECK(host_out && host_in, JETS_ERR_INVALID, "null host pointer");
    JETS_CHECK(D->kind == 0, JETS_ERR_UNSUPPORTED, "the host-buffer pipeline is implemented for block-banded operators");
    JETS_CHECK(D->A->linear, JETS_ERR_NOT_LINEAR, "A'*(A*m) requires a linear operator");
    if (!D->pipe || (nchunks > 0 && D->pipe->K != std::max(1, std::min<int>(nchunks, D->nloc / std::max(1, D->halo))))) {
      D->pipe.reset();
      make_pipe(D, nchunks > 0 ? nchunks : 32);
    }
    pipe_step(D, reinterpret_cast<char*>(host_out), reinterpret_cast<const char*>(host_in));
  });
}

int jets_dist_op_join(jets_dist_op D) {
  return guard([&] {
    require_ready();
    JETS_CHECK(D && live_ops().count(D), JETS_ERR_INVALID, "null or destroyed distributed operator handle");
    if (D->pipe && D->pipe->have_prev_step) {
      CUDA_TRY(cudaStreamWaitEvent(ctx().stream, D->pipe->ev_down, 0));
      CUDA_TRY(cudaStreamWaitEvent(ctx().stream, D->pipe->ev_comp, 0));
    }
  });
}

int32_t jets_dist_pipeline_schedule(int32_t nloc, int32_t halo, int32_t nchunks, int32_t has_prev, int32_t has_next, const uint8_t* nz,
                                    int32_t cap, int32_t* items, int32_t* chunk_bounds, int32_t* up_need, int32_t* nchunks_out) {
  int32_t count = -1;
  guard([&] {
    JETS_CHECK(nloc >= 1 && halo >= 0 && halo <= kMaxHalo && nloc >= halo && nz && items && chunk_bounds && up_need && nchunks_out && cap >= 1,
               JETS_ERR_INVALID, "bad arguments");
    HostPipe P;
    const int C = nloc + 2 * halo;
    pipe_schedule(P, nloc, halo, nchunks, has_prev != 0, has_next != 0, [&](int r, int j) { return nz[(size_t)r * C + j] != 0; });
    JETS_CHECK((int)P.seq.size() <= cap && P.K <= cap, JETS_ERR_INVALID, "schedule has %d items, capacity %d", (int)P.seq.size(), cap);
    for (size_t i = 0; i < P.seq.size(); ++i) { items[2 * i] = P.seq[i].what; items[2 * i + 1] = P.seq[i].k; }
    for (int k = 0; k < P.K; ++k) { chunk_bounds[2 * k] = P.chunks[k].first; chunk_bounds[2 * k + 1] = P.chunks[k].second; up_need[k] = P.up_need[k]; }
    *nchunks_out = P.K;
    count = (int32_t)P.seq.size();
  });
  return count;
}

int32_t jets_dist_op_info(jets_dist_op D, int32_t what) {
  if (!D || !live_ops().count(D)) return -1;
  switch (what) {
    case 0: return D->nloc;
    case 1: return D->halo;
    case 2: return (D->has_prev ? 1 : 0) | (D->has_next ? 2 : 0);
    case 3: return D->pipe ? D->pipe->K : 0;
    case 4: {   // launches of one monolithic apply
      int n = 0;
      for (auto& p : D->mono)
        if (p) n = std::max<int>(n, (int)p->steps.size());
      return n;
    }
    case 5: return D->kind;
    case 6: {   // work units that gave up waiting for a neighbour (results are invalid if != 0); synchronises
      if (!D->arena) return 0;
      uint32_t v = 0;
      if (cudaDeviceSynchronize() != cudaSuccess) return -1;
      if (cudaMemcpy(&v, my_flags_base(D) + 2 * kGateFlags * kGateFlagStride, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
      return (int32_t)v;
    }
    default: return -1;
  }
}

}  // extern "C"
