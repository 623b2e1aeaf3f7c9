// C ABI of libjets_b200.so (include/jets_b200.h): context, device storage, operator trees,
// linearization, apply, reductions.  No exceptions leave this file.
#include <cstdarg>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include "common.hpp"

namespace jets {

extern uint64_t g_epoch;
std::shared_ptr<Plan> build_plan(jets_op a, int mode, int accumulate, bool io_ok, int engine);

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

Context& ctx() {
  static Context c;
  return c;
}
void require_ready() {
  JETS_CHECK(ctx().ready, JETS_ERR_CUDA, "jets_init() has not been called (or failed): no CUDA context");
}

Storage::~Storage() {
  if (alloc) cudaFree(alloc);
}

static jets_buf make_buf(int dt, std::shared_ptr<Storage> st, int64_t off, int32_t nblocks,
                         const int64_t* len) {
  auto* b = new jets_buf_s();
  b->dtype = dt;
  b->st = std::move(st);
  b->off = off;
  b->blk_off.assign(1, 0);
  for (int i = 0; i < nblocks; ++i) b->blk_off.push_back(b->blk_off.back() + len[i]);
  return b;
}

static void check_buf(jets_buf x) { JETS_CHECK(x && x->refs > 0, JETS_ERR_INVALID, "null or destroyed buffer handle"); }
static void check_op(jets_op a) { JETS_CHECK(a && a->refs > 0, JETS_ERR_INVALID, "null or destroyed operator handle"); }
static void check_dt(int dt) { JETS_CHECK(dt >= JETS_F32 && dt <= JETS_C128, JETS_ERR_DTYPE, "unsupported dtype %d", dt); }
static void check_real(int dt, const char* what) {
  JETS_CHECK(!is_cplx(dt), JETS_ERR_UNSUPPORTED, "%s is not implemented for complex eltypes", what);
}

static void buf_release(jets_buf x) {
  if (x && --x->refs == 0) delete x;
}
static void op_release(jets_op a) {
  if (a && --a->refs == 0) delete a;
}

static jets_buf alloc_buf(int dt, int32_t nblocks, const int64_t* len) {
  int64_t total = 0;
  for (int i = 0; i < nblocks; ++i) {
    JETS_CHECK(len[i] >= 0, JETS_ERR_SHAPE, "negative block length");
    total += len[i];
  }
  auto st = std::make_shared<Storage>();
  st->bytes = (size_t)total * dsize(dt);
  const size_t padded = ((st->bytes + 255) & ~(size_t)255) + 2 * kGuardBytes;
  CUDA_TRY(cudaMalloc(&st->alloc, padded));
  CUDA_TRY(cudaMemsetAsync(st->alloc, 0, padded, ctx().stream));
  st->data = reinterpret_cast<char*>(st->alloc) + kGuardBytes;
  st->guarded = true;
  return make_buf(dt, st, 0, nblocks, len);
}

static bool buf_io_ok(jets_buf x) {
  return x->guarded() && (reinterpret_cast<uintptr_t>(x->ptr()) & 15) == 0;
}

// ------------------------------------------------------------------ op helpers -----------
static jets_op new_op(Kind k, int dt) {
  auto* a = new jets_op_s();
  a->kind = k;
  a->dtype = dt;
  return a;
}
static Space space1(int64_t n) {
  Space s;
  s.len = {n};
  s.is_block = false;
  return s;
}
static bool is_wrapper(jets_op a) { return a->kind == K_LNVIEW || a->kind == K_ADJ; }
static jets_op strip(jets_op a) {
  while (is_wrapper(a)) a = a->kids[0];
  return a;
}
static jets_op retain(jets_op a) {
  a->refs++;
  return a;
}
static jets_op as_linear(jets_op a) {  // returns a retained handle
  if (a->linear) return retain(a);
  jets_op v = new_op(K_LNVIEW, a->dtype);
  v->dom = a->dom; v->rng = a->rng; v->linear = true;
  v->kids.push_back(retain(a));
  return v;
}
static jets_op adjoint_of(jets_op a) {  // returns a retained handle
  JETS_CHECK(a->linear, JETS_ERR_NOT_LINEAR,
             "adjoint is defined for linear operators only (src/Jets.jl:382-392): linearize with jacobian first");
  if (a->kind == K_ADJ) return retain(a->kids[0]);
  jets_op v = new_op(K_ADJ, a->dtype);
  v->dom = a->rng; v->rng = a->dom; v->linear = true;
  v->kids.push_back(retain(a));
  return v;
}
static bool needs_point(jets_op a) {
  if (a->kind == K_PW) return true;
  for (jets_op k : a->kids)
    if (needs_point(k)) return true;
  return false;
}

static void apply_impl(jets_op a, int mode, jets_buf out, jets_buf in, int accumulate, const ApplyCoef* coef = nullptr,
                       double* norm_out = nullptr) {
  require_ready();
  check_op(a); check_buf(out); check_buf(in);
  JETS_CHECK(mode >= 0 && mode <= 2, JETS_ERR_INVALID, "bad mode %d", mode);
  JETS_CHECK(in->dtype == a->dtype && out->dtype == a->dtype, JETS_ERR_DTYPE,
             "operator eltype %d but in/out eltypes %d/%d", a->dtype, in->dtype, out->dtype);
  if (mode == JETS_MODE_DFT)
    JETS_CHECK(a->linear, JETS_ERR_NOT_LINEAR, "mul!(m, A', d) requires a linear operator (src/Jets.jl:392)");
  const Space& isp = mode == JETS_MODE_DFT ? a->rng : a->dom;
  const Space& osp = mode == JETS_MODE_DFT ? a->dom : a->rng;
  JETS_CHECK(in->length() == isp.total(), JETS_ERR_SHAPE, "input has %lld elements, operator expects %lld",
             (long long)in->length(), (long long)isp.total());
  JETS_CHECK(out->length() == osp.total(), JETS_ERR_SHAPE, "output has %lld elements, operator produces %lld",
             (long long)out->length(), (long long)osp.total());
  // in/out may not overlap (the reference would also produce garbage): cheap check on bases
  const bool io_ok = buf_io_ok(in) && buf_io_ok(out);
  const int engine = ctx().fused_engine;
  const int key = mode | ((accumulate ? 1 : 0) << 2) | ((io_ok ? 1 : 0) << 3) | (engine << 4);
  std::shared_ptr<Plan> plan;
  auto it = a->plans.find(key);
  if (it != a->plans.end() && it->second->valid()) plan = it->second;
  else {
    for (auto p = a->plans.begin(); p != a->plans.end();)      // stale plans hold the old linearization points: drop them now
      p = p->second->valid() ? std::next(p) : a->plans.erase(p);
    plan = build_plan(a, mode, accumulate, io_ok, engine);
    a->plans[key] = plan;
  }
  if (ctx().capturing) {        // the graph replays raw pointers into the plan's tables and the operator's state
    ctx().capture_keep.push_back(plan);
    a->refs++;
    ctx().capture_keep.push_back(std::shared_ptr<void>(a, [](void* p) { op_release(reinterpret_cast<jets_op>(p)); }));
  }
  if (coef) {
    check_real(a->dtype, "jets_apply_axpby");
    // out = cA*(A in) + cO*out: in the kernel's store epilogue when the apply is one bundle launch ...
    if (run_plan_axpby(*plan, a->dtype, in->ptr(), out->ptr(), *coef)) {
      // the norm right behind the launch that wrote the vector: for solver-sized vectors it is served from L2
      if (norm_out) vec_reduce(a->dtype, 1, out->ptr(), nullptr, out->length(), 2.0, norm_out, ctx().stream);
      return;
    }
    // ... else through a temporary owned by the operator (dense / staged plans)
    const size_t bytes = (size_t)out->length() * dsize(a->dtype);
    if (!a->axpby_tmp || a->axpby_tmp_bytes < bytes) {
      if (a->axpby_tmp) { cudaStreamSynchronize(ctx().stream); cudaFree(a->axpby_tmp); }
      CUDA_TRY(cudaMalloc(&a->axpby_tmp, bytes + 2 * kGuardBytes));
      a->axpby_tmp_bytes = bytes;
    }
    char* tmp = reinterpret_cast<char*>(a->axpby_tmp) + kGuardBytes;
    // the plan was built for guarded, aligned in/out; the temporary is both
    run_plan(*plan, a->dtype, in->ptr(), tmp);
    vec_axpby_dev(a->dtype, out->ptr(), out->length(), coef->a_ptr, coef->a_const, coef->a_flags, tmp, coef->o_ptr,
                  coef->o_const, coef->o_flags, out->ptr(), ctx().stream);
    if (norm_out) vec_reduce(a->dtype, 1, out->ptr(), nullptr, out->length(), 2.0, norm_out, ctx().stream);   // a pass of its own
    return;
  }
  run_plan(*plan, a->dtype, in->ptr(), out->ptr());
}

static void set_point_impl(jets_op a, jets_buf mo) {
  switch (a->kind) {
    case K_PW: {
      JETS_CHECK(mo->dtype == a->dtype, JETS_ERR_DTYPE, "point eltype mismatch");
      JETS_CHECK(mo->length() == a->dom.total(), JETS_ERR_SHAPE, "point has %lld elements, domain %lld",
                 (long long)mo->length(), (long long)a->dom.total());
      mo->refs++;
      buf_release(a->mo);
      a->mo = mo;  // by reference (src/Jets.jl:298)
      return;
    }
    case K_LNVIEW: case K_ADJ: set_point_impl(a->kids[0], mo); return;
    case K_COMPOSE: {  // src/Jets.jl:578-589
      const int n = (int)a->kids.size();
      int lowest = -1;  // lowest index that still needs a point
      for (int i = 0; i < n; ++i)
        if (needs_point(a->kids[i])) { lowest = i; break; }
      if (lowest < 0) return;
      int64_t len = mo->length();
      jets_buf m = alloc_buf(mo->dtype, 1, &len);  // _m = copy(mo)
      CUDA_TRY(cudaMemcpyAsync(m->ptr(), mo->ptr(), (size_t)len * dsize(mo->dtype), cudaMemcpyDeviceToDevice, ctx().stream));
      for (int i = n - 1; i >= lowest; --i) {
        if (needs_point(a->kids[i])) set_point_impl(a->kids[i], m);
        if (i > lowest) {
          int64_t rl = a->kids[i]->rng.total();
          jets_buf nx = alloc_buf(mo->dtype, 1, &rl);
          apply_impl(a->kids[i], JETS_MODE_F, nx, m, 0);  // _m = ops[i] * _m
          buf_release(m);
          m = nx;
        }
      }
      buf_release(m);
      return;
    }
    case K_SUM:  // :710-715
      for (jets_op k : a->kids) set_point_impl(k, mo);
      return;
    case K_BLOCK: {  // :1059-1066
      for (int c = 0; c < a->C; ++c) {
        jets_buf mc = mo;
        bool view = false;
        if (a->dom.is_block || a->C > 1) {
          // getblock(mo, c): a view of block c
          int64_t off = 0;
          for (int k = 0; k < c; ++k) off += a->dom.len[k];
          int64_t l = a->dom.len[c];
          mc = make_buf(mo->dtype, mo->st, mo->off + off, 1, &l);
          view = true;
        }
        for (int r = 0; r < a->R; ++r) {
          jets_op kid = a->kids[r + (size_t)c * a->R];
          if (needs_point(kid)) set_point_impl(kid, mc);
        }
        if (view) buf_release(mc);
      }
      return;
    }
    default: return;  // linear leaves ignore the point
  }
}

static jets_op clone_tree(jets_op a) {  // copy(F,false): new nodes, shared (immutable) state buffers
  jets_op c = new_op(a->kind, a->dtype);
  c->dom = a->dom; c->rng = a->rng; c->linear = a->linear;
  c->w = a->w;
  if (c->w) c->w->refs++;
  c->a = a->a; c->ai = a->ai; c->p = a->p; c->fn = a->fn;
  c->gidx = a->gidx; c->gidx64 = a->gidx64;
  c->rows = a->rows; c->cols = a->cols; c->nrhs = a->nrhs;
  c->mo = a->mo;
  if (c->mo) c->mo->refs++;
  c->sgn = a->sgn; c->R = a->R; c->C = a->C;
  for (jets_op k : a->kids) c->kids.push_back(clone_tree(k));
  return c;
}

}  // namespace jets

jets_op_s::~jets_op_s() {
  plans.clear();
  if (axpby_tmp) cudaFree(axpby_tmp);
  if (w && --w->refs == 0) delete w;
  if (mo && --mo->refs == 0) delete mo;
  for (jets_op k : kids)
    if (k && --k->refs == 0) delete k;
}

using namespace jets;

extern "C" {

int jets_abi_version(void) { return JETS_B200_ABI_VERSION; }
const char* jets_last_error(void) { return g_err; }

int jets_init(int device) {
  return guard([&] {
    Context& c = ctx();
    if (c.ready && c.device == device) return;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    JETS_CHECK(e == cudaSuccess && n > 0, JETS_ERR_CUDA,
               "no CUDA device available (%s); libjets_b200 has no CPU fallback", cudaGetErrorString(e));
    JETS_CHECK(device >= 0 && device < n, JETS_ERR_INVALID, "device %d out of range [0,%d)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device));
    JETS_CHECK(p.major == 10, JETS_ERR_CUDA, "device %d is sm_%d%d; libjets_b200 is built for sm_100a only",
               device, p.major, p.minor);
    c.device = device;
    c.sm_count = p.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    c.own_stream = true;
    CUDA_TRY(cudaMallocHost(&c.host_scratch, 64 * sizeof(double)));
    c.dev_scratch_elems = (size_t)c.sm_count * 8 + 64;
    CUDA_TRY(cudaMalloc(&c.dev_scratch, (c.dev_scratch_elems + 64) * sizeof(double)));
    CUDA_TRY(cudaMemset(c.dev_scratch, 0, (c.dev_scratch_elems + 64) * sizeof(double)));   // result slots; slot 60 = the reductions' block ticket
    if (const char* v = getenv("JETS_B200_FAST_VARIANT")) c.fast_variant = atoi(v);
    if (const char* v = getenv("JETS_B200_NO_FAST")) c.no_fast = atoi(v);
    if (const char* v = getenv("JETS_B200_NO_BUNDLE")) c.no_bundle = atoi(v);
    if (const char* v = getenv("JETS_B200_BUNDLE_NX")) c.bundle_nx = atoi(v);
    if (const char* v = getenv("JETS_B200_BUNDLE_NS")) c.bundle_ns = atoi(v);
    if (const char* v = getenv("JETS_B200_BUNDLE_BMAX")) c.bundle_bmax = atoi(v);
    if (const char* v = getenv("JETS_B200_NO_PDL")) c.no_pdl = atoi(v);
    if (const char* v = getenv("JETS_B200_VEC_PDL")) c.vec_pdl = atoi(v);
    if (const char* v = getenv("JETS_B200_STATIC_SCHED")) c.static_sched = atoi(v);
    if (const char* v = getenv("JETS_B200_GRID")) c.grid_limit = atoi(v);
    if (const char* v = getenv("JETS_B200_DIST_EARLY_CTAS")) c.dist_early_ctas = atoi(v);
    if (const char* v = getenv("JETS_B200_NO_PRE_STATE")) c.no_pre_state = atoi(v);
    if (const char* v = getenv("JETS_B200_GROUP_STREAMS")) c.group_streams = atoi(v);
    if (const char* v = getenv("JETS_B200_NO_TAIL_SPLIT")) c.no_tail_split = atoi(v);
    if (const char* v = getenv("JETS_B200_TAIL_MIN_UNITS")) c.tail_min_units = atoll(v);
    if (const char* v = getenv("JETS_B200_TILE_ELEMS")) c.tile_elems = atoll(v);
    if (const char* v = getenv("JETS_B200_TAIL_MIN_ROWS")) c.tail_min_rows = std::max(2, atoi(v));
    if (const char* v = getenv("JETS_B200_TAIL_DIV")) c.tail_div = std::max(2, atoi(v));
    if (const char* v = getenv("JETS_B200_TAIL_SUB_MIN")) c.tail_sub_min = std::max(1, atoi(v));
    if (const char* v = getenv("JETS_B200_NO_FIRST_STATIC")) c.no_first_static = atoi(v);
    if (const char* v = getenv("JETS_B200_TRACE")) {
      if (atoi(v)) {
        const size_t n = (size_t)Context::kTraceLaunches * Context::kTraceCtas * 8 * sizeof(unsigned long long);
        CUDA_TRY(cudaMalloc(&c.trace_buf, n));
        CUDA_TRY(cudaMemset(c.trace_buf, 0, n));
      }
    }
    c.ready = true;
  });
}

int jets_shutdown(void) {
  return guard([&] {
    Context& c = ctx();
    if (!c.ready) return;
    cudaStreamSynchronize(c.stream);
    if (c.own_stream) cudaStreamDestroy(c.stream);
    cudaFreeHost(c.host_scratch);
    cudaFree(c.dev_scratch);
    c = Context();
  });
}

int jets_stream_set(void* s) {
  return guard([&] {
    require_ready();
    Context& c = ctx();
    JETS_CHECK(!c.on_aux, JETS_ERR_INVALID, "jets_stream_set while forked onto an auxiliary stream");
    if (c.own_stream && c.stream) { cudaStreamSynchronize(c.stream); cudaStreamDestroy(c.stream); }
    c.stream = reinterpret_cast<cudaStream_t>(s);
    c.own_stream = false;
  });
}
void* jets_stream_get(void) { return ctx().stream; }
int jets_stream_fork(int aux) {
  return guard([&] {
    require_ready();
    Context& c = ctx();
    JETS_CHECK(aux >= 0 && aux < 2, JETS_ERR_INVALID, "auxiliary stream index must be 0 or 1");
    if (!c.aux[aux]) {
      int lo = 0, hi = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_TRY(cudaStreamCreateWithPriority(&c.aux[aux], cudaStreamNonBlocking, hi));   // hi = greatest priority
      CUDA_TRY(cudaEventCreateWithFlags(&c.ev_join[aux], cudaEventDisableTiming));
    }
    if (!c.ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    if (!c.on_aux) c.main_stream = c.stream;
    CUDA_TRY(cudaEventRecord(c.ev_fork, c.main_stream));
    CUDA_TRY(cudaStreamWaitEvent(c.aux[aux], c.ev_fork, 0));
    c.stream = c.aux[aux];
    c.on_aux = true;
  });
}
int jets_stream_main(void) {
  return guard([&] {
    require_ready();
    Context& c = ctx();
    if (c.on_aux) { c.stream = c.main_stream; c.on_aux = false; }
  });
}
int jets_stream_join(int aux) {
  return guard([&] {
    require_ready();
    Context& c = ctx();
    JETS_CHECK(aux >= 0 && aux < 2 && c.aux[aux], JETS_ERR_INVALID, "jets_stream_join(%d) without a fork", aux);
    if (c.on_aux) { c.stream = c.main_stream; c.on_aux = false; }
    CUDA_TRY(cudaEventRecord(c.ev_join[aux], c.aux[aux]));
    CUDA_TRY(cudaStreamWaitEvent(c.stream, c.ev_join[aux], 0));
  });
}
int jets_sync(void) {
  return guard([&] { require_ready(); CUDA_TRY(cudaStreamSynchronize(ctx().stream)); });
}
int64_t jets_launch_count(void) { return ctx().launches; }
int64_t jets_debug_trace(uint64_t* host, int64_t capacity) {
  // JETS_B200_TRACE=1: copies the ring of per-CTA launch timelines (kTraceLaunches x kTraceCtas x 8 globaltimer
  // stamps) to `host`; returns the number of bundle launches traced so far, -1 when tracing is off.
  Context& c = ctx();
  if (!c.trace_buf) return -1;
  const int64_t n = (int64_t)Context::kTraceLaunches * Context::kTraceCtas * 8;
  if (!host || capacity < n) return -1;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpy(host, c.trace_buf, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return c.trace_count;
}
int jets_device_sm_count(void) { return ctx().sm_count; }
int jets_set_fused_engine(int which) {
  return guard([&] {
    JETS_CHECK(which >= 0 && which <= 3, JETS_ERR_INVALID, "engine must be 0,1,2,3");
    ctx().fused_engine = which;
  });
}

// ------------------------------------------------------------------ buffers --------------
int jets_buf_create(jets_dtype dt, int32_t nblocks, const int64_t* block_len, jets_buf* out) {
  return guard([&] {
    require_ready(); check_dt(dt);
    JETS_CHECK(out && block_len && nblocks >= 1, JETS_ERR_INVALID, "bad arguments");
    *out = alloc_buf(dt, nblocks, block_len);
  });
}
int jets_buf_wrap(jets_dtype dt, void* devptr, int32_t nblocks, const int64_t* block_len, jets_buf* out) {
  return guard([&] {
    require_ready(); check_dt(dt);
    JETS_CHECK(out && block_len && nblocks >= 1 && devptr, JETS_ERR_INVALID, "bad arguments");
    auto st = std::make_shared<Storage>();
    st->data = reinterpret_cast<char*>(devptr);
    int64_t total = 0;
    for (int i = 0; i < nblocks; ++i) total += block_len[i];
    st->bytes = (size_t)total * dsize(dt);
    *out = make_buf(dt, st, 0, nblocks, block_len);
  });
}
int jets_buf_view(jets_buf x, int32_t first, int32_t n, jets_buf* out) {
  return guard([&] {
    check_buf(x);
    JETS_CHECK(out && first >= 0 && n >= 1 && first + n <= x->nblocks(), JETS_ERR_INVALID,
               "block range [%d,%d) outside 0..%d", first, first + n, x->nblocks());
    std::vector<int64_t> len(n);
    for (int i = 0; i < n; ++i) len[i] = x->blk_off[first + i + 1] - x->blk_off[first + i];
    *out = make_buf(x->dtype, x->st, x->off + x->blk_off[first], n, len.data());
  });
}
int jets_buf_reshape(jets_buf x, int32_t nblocks, const int64_t* block_len, jets_buf* out) {
  return guard([&] {
    check_buf(x);
    JETS_CHECK(out && block_len && nblocks >= 1, JETS_ERR_INVALID, "bad arguments");
    int64_t total = 0;
    for (int i = 0; i < nblocks; ++i) total += block_len[i];
    JETS_CHECK(total == x->length(), JETS_ERR_SHAPE, "dimension mismatch, unable to reshape block array");
    *out = make_buf(x->dtype, x->st, x->off, nblocks, block_len);
  });
}
int jets_buf_retain(jets_buf x) { return guard([&] { check_buf(x); x->refs++; }); }
int jets_buf_destroy(jets_buf x) { return guard([&] { check_buf(x); buf_release(x); }); }
int jets_buf_dtype(jets_buf x) { return x ? x->dtype : -1; }
int32_t jets_buf_nblocks(jets_buf x) { return x ? x->nblocks() : -1; }
int64_t jets_buf_length(jets_buf x) { return x ? x->length() : -1; }
int jets_buf_block_range(jets_buf x, int32_t b, int64_t* first1, int64_t* last1) {
  return guard([&] {
    check_buf(x);
    JETS_CHECK(b >= 0 && b < x->nblocks(), JETS_ERR_INVALID, "block %d out of range", b);
    *first1 = x->blk_off[b] + 1;   // 1-based inclusive, src/Jets.jl:745-747
    *last1 = x->blk_off[b + 1];
  });
}
void* jets_buf_devptr(jets_buf x) { return x ? x->ptr() : nullptr; }

static void xfer(jets_buf x, int32_t block, void* host, int64_t count, bool up, bool async) {
  require_ready(); check_buf(x);
  JETS_CHECK(host || count == 0, JETS_ERR_INVALID, "null host pointer");
  char* p;
  int64_t n;
  if (block < 0) { p = x->ptr(); n = x->length(); }
  else {
    JETS_CHECK(block < x->nblocks(), JETS_ERR_INVALID, "block %d out of range", block);
    p = x->block_ptr(block);
    n = x->blk_off[block + 1] - x->blk_off[block];
  }
  JETS_CHECK(count == n, JETS_ERR_SHAPE, "host array has %lld elements, device block has %lld", (long long)count, (long long)n);
  const size_t bytes = (size_t)n * dsize(x->dtype);
  if (up) CUDA_TRY(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, ctx().stream));
  else CUDA_TRY(cudaMemcpyAsync(host, p, bytes, cudaMemcpyDeviceToHost, ctx().stream));
  if (!async) CUDA_TRY(cudaStreamSynchronize(ctx().stream));
}
int jets_buf_upload(jets_buf x, int32_t b, const void* h, int64_t n) { return guard([&] { xfer(x, b, const_cast<void*>(h), n, true, false); }); }
int jets_buf_download(jets_buf x, int32_t b, void* h, int64_t n) { return guard([&] { xfer(x, b, h, n, false, false); }); }
int jets_buf_upload_async(jets_buf x, int32_t b, const void* h, int64_t n) { return guard([&] { xfer(x, b, const_cast<void*>(h), n, true, true); }); }
int jets_buf_download_async(jets_buf x, int32_t b, void* h, int64_t n) { return guard([&] { xfer(x, b, h, n, false, true); }); }

// x[first .. first+count) <-> host: scalar getindex/setindex! and SymmetricArray element access
static void xfer_range(jets_buf x, int64_t first, void* host, int64_t count, bool up) {
  require_ready(); check_buf(x);
  JETS_CHECK(host || count == 0, JETS_ERR_INVALID, "null host pointer");
  JETS_CHECK(first >= 0 && count >= 0 && first + count <= x->length(), JETS_ERR_SHAPE,
             "range [%lld, %lld) outside a vector of %lld elements", (long long)first, (long long)(first + count), (long long)x->length());
  const size_t w = dsize(x->dtype);
  char* p = x->ptr() + (size_t)first * w;
  if (up) CUDA_TRY(cudaMemcpyAsync(p, host, (size_t)count * w, cudaMemcpyHostToDevice, ctx().stream));
  else CUDA_TRY(cudaMemcpyAsync(host, p, (size_t)count * w, cudaMemcpyDeviceToHost, ctx().stream));
  CUDA_TRY(cudaStreamSynchronize(ctx().stream));
}
int jets_buf_write(jets_buf x, int64_t first, const void* h, int64_t n) { return guard([&] { xfer_range(x, first, const_cast<void*>(h), n, true); }); }
int jets_buf_read(jets_buf x, int64_t first, void* h, int64_t n) { return guard([&] { xfer_range(x, first, h, n, false); }); }

int jets_buf_copy(jets_buf dst, jets_buf src) {
  return guard([&] {
    require_ready(); check_buf(dst); check_buf(src);
    JETS_CHECK(dst->dtype == src->dtype, JETS_ERR_DTYPE, "copy between different eltypes");
    JETS_CHECK(dst->length() == src->length(), JETS_ERR_SHAPE, "copy: %lld vs %lld elements", (long long)dst->length(), (long long)src->length());
    CUDA_TRY(cudaMemcpyAsync(dst->ptr(), src->ptr(), (size_t)dst->length() * dsize(dst->dtype), cudaMemcpyDeviceToDevice, ctx().stream));
  });
}
int jets_buf_fill(jets_buf x, double a) { return jets_buf_fill_c(x, a, 0.0); }
int jets_buf_fill_c(jets_buf x, double re, double im) {
  return guard([&] {
    require_ready(); check_buf(x);
    if (is_cplx(x->dtype)) cvec_fill(x->dtype, x->ptr(), x->length(), re, im, ctx().stream);
    else {
      JETS_CHECK(im == 0.0, JETS_ERR_DTYPE, "complex fill value for a real vector");
      vec_fill(x->dtype, x->ptr(), x->length(), re, ctx().stream);
    }
  });
}
int jets_buf_rand(jets_buf x, uint64_t seed, uint64_t off, int dist) {
  return guard([&] {
    require_ready(); check_buf(x);
    JETS_CHECK(dist == 0 || dist == 1, JETS_ERR_INVALID, "dist must be 0 (uniform) or 1 (normal)");
    if (is_cplx(x->dtype)) {
      // rand(ComplexF64): both parts U[0,1); randn: both parts N(0, 1/2).  Part k of element i draws
      // from counter 2i+k, so any partition of the logical vector sees the same numbers.
      const int rdt = real_of(x->dtype);
      vec_rand(rdt, x->ptr(), 2 * x->length(), seed, 2 * off, dist, ctx().stream);
      if (dist == 1) {
        const double c = 0.70710678118654752440;
        const void* px = x->ptr();
        vec_lincomb(rdt, x->ptr(), 2 * x->length(), 1, &c, &px, ctx().stream);
      }
    } else {
      vec_rand(x->dtype, x->ptr(), x->length(), seed, off, dist, ctx().stream);
    }
  });
}

// ------------------------------------------------------------------ reductions -----------
static void check_pair(jets_buf x, jets_buf y) {
  check_buf(x); check_buf(y);
  JETS_CHECK(x->dtype == y->dtype, JETS_ERR_DTYPE, "eltype mismatch");
  JETS_CHECK(x->length() == y->length(), JETS_ERR_SHAPE, "length mismatch %lld vs %lld", (long long)x->length(), (long long)y->length());
}
static double* result_slot(int i) { return ctx().dev_scratch + ctx().dev_scratch_elems + i; }
static void fetch(int n, double* out) {
  Context& c = ctx();
  CUDA_TRY(cudaMemcpyAsync(c.host_scratch, result_slot(0), n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  for (int i = 0; i < n; ++i) out[i] = c.host_scratch[i];
}
static void norm_to(jets_buf x, double p, double* dev_out) {
  int kind;
  if (p == 2.0) kind = 1;
  else if (p == 1.0) kind = 2;
  else if (p == 0.0) kind = 3;
  else if (std::isinf(p) && p > 0) kind = 4;
  else if (std::isinf(p) && p < 0) kind = 5;
  else kind = 6;
  if (is_cplx(x->dtype)) {
    // |z| replaces |x|; kinds 1..6 of vec_reduce map onto 2..7 of cvec_reduce
    const int finish = kind == 1 ? 1 : kind == 6 ? 2 : 0;
    cvec_reduce(x->dtype, kind + 1, x->ptr(), nullptr, nullptr, x->length(), p, finish, dev_out, ctx().stream);
    return;
  }
  vec_reduce(x->dtype, kind, x->ptr(), nullptr, x->length(), p, dev_out, ctx().stream);
}
int jets_dot(jets_buf x, jets_buf y, double* out) {
  return guard([&] {
    require_ready(); check_pair(x, y);
    if (is_cplx(x->dtype)) cvec_reduce(x->dtype, 0, x->ptr(), y->ptr(), nullptr, x->length(), 0, 0, result_slot(0), ctx().stream);
    else vec_reduce(x->dtype, 0, x->ptr(), y->ptr(), x->length(), 0, result_slot(0), ctx().stream);
    fetch(1, out);
  });
}
int jets_dot_c(jets_buf x, jets_buf y, double* out_re_im) {
  return guard([&] {
    require_ready(); check_pair(x, y);
    JETS_CHECK(out_re_im, JETS_ERR_INVALID, "null out");
    if (is_cplx(x->dtype)) {
      cvec_reduce(x->dtype, 0, x->ptr(), y->ptr(), nullptr, x->length(), 0, 0, result_slot(0), ctx().stream);
      cvec_reduce(x->dtype, 1, x->ptr(), y->ptr(), nullptr, x->length(), 0, 0, result_slot(1), ctx().stream);
      fetch(2, out_re_im);
    } else {
      vec_reduce(x->dtype, 0, x->ptr(), y->ptr(), x->length(), 0, result_slot(0), ctx().stream);
      fetch(1, out_re_im);
      out_re_im[1] = 0.0;
    }
  });
}
int jets_norm(jets_buf x, double p, double* out) {
  return guard([&] { require_ready(); check_buf(x); norm_to(x, p, result_slot(0)); fetch(1, out); });
}
int jets_norm_weighted(jets_buf x, jets_buf w, double p, double* out) {
  return guard([&] {
    require_ready(); check_buf(x); check_buf(w);
    JETS_CHECK(w->dtype == JETS_F64 && w->length() == x->length(), JETS_ERR_SHAPE, "weights must be Float64, one per element");
    const double* wp = reinterpret_cast<const double*>(w->ptr());
    int kind, finish = 0;
    if (p == 2.0) { kind = 2; finish = 1; }
    else if (p == 1.0) kind = 3;
    else if (p == 0.0) kind = 4;
    else if (std::isinf(p) && p > 0) kind = 5;
    else if (std::isinf(p) && p < 0) kind = 6;
    else { kind = 7; finish = 2; }
    cvec_reduce(x->dtype, kind, x->ptr(), nullptr, wp, x->length(), p, finish, result_slot(0), ctx().stream);
    fetch(1, out);
  });
}
int jets_abs(jets_buf out, jets_buf x) {
  return guard([&] {
    require_ready(); check_buf(out); check_buf(x);
    JETS_CHECK(is_cplx(x->dtype) && out->dtype == real_of(x->dtype), JETS_ERR_DTYPE,
               "abs.(x): x must be complex and out its real eltype");
    JETS_CHECK(out->length() == x->length(), JETS_ERR_SHAPE, "length mismatch");
    cvec_abs(x->dtype, out->ptr(), x->ptr(), x->length(), ctx().stream);
  });
}
int jets_extrema(jets_buf x, double* mn, double* mx) {
  return guard([&] {
    require_ready(); check_buf(x);
    JETS_CHECK(x->length() > 0, JETS_ERR_SHAPE, "extrema of an empty vector");
    check_real(x->dtype, "extrema (complex numbers are not ordered)");
    vec_reduce(x->dtype, 7, x->ptr(), nullptr, x->length(), 0, result_slot(0), ctx().stream);
    vec_reduce(x->dtype, 8, x->ptr(), nullptr, x->length(), 0, result_slot(1), ctx().stream);
    double r[2];
    fetch(2, r);
    *mn = r[0]; *mx = r[1];
  });
}
int jets_lincomb(jets_buf out, int32_t n, const double* c, const jets_buf* x) {
  return guard([&] {
    require_ready(); check_buf(out);
    JETS_CHECK(n >= 1 && n <= 4 && c && x, JETS_ERR_INVALID, "lincomb takes 1..4 terms");
    const void* px[4];
    for (int i = 0; i < n; ++i) { check_pair(out, x[i]); px[i] = x[i]->ptr(); }
    if (is_cplx(out->dtype)) vec_lincomb(real_of(out->dtype), out->ptr(), 2 * out->length(), n, c, px, ctx().stream);
    else vec_lincomb(out->dtype, out->ptr(), out->length(), n, c, px, ctx().stream);
  });
}
int jets_lincomb_c(jets_buf out, int32_t n, const double* c_re_im, const jets_buf* x) {
  return guard([&] {
    require_ready(); check_buf(out);
    JETS_CHECK(n >= 1 && n <= 4 && c_re_im && x, JETS_ERR_INVALID, "lincomb takes 1..4 terms");
    JETS_CHECK(is_cplx(out->dtype), JETS_ERR_DTYPE, "complex coefficients need a complex vector");
    const void* px[4];
    for (int i = 0; i < n; ++i) { check_pair(out, x[i]); px[i] = x[i]->ptr(); }
    cvec_lincomb(out->dtype, out->ptr(), out->length(), n, c_re_im, px, ctx().stream);
  });
}
int jets_hadamard(jets_buf out, jets_buf x, jets_buf y) {
  return guard([&] {
    require_ready(); check_pair(out, x); check_pair(out, y);
    if (is_cplx(out->dtype)) cvec_hadamard(out->dtype, out->ptr(), x->ptr(), y->ptr(), out->length(), 0, ctx().stream);
    else vec_hadamard(out->dtype, out->ptr(), x->ptr(), y->ptr(), out->length(), ctx().stream);
  });
}

// ------------------------------------------------------------------ device scalars -------
int jets_scalar_create(jets_scalar* out) {
  return guard([&] {
    require_ready();
    auto* s = new jets_scalar_s();
    CUDA_TRY(cudaMalloc(&s->dev, sizeof(double)));
    CUDA_TRY(cudaMemsetAsync(s->dev, 0, sizeof(double), ctx().stream));
    *out = s;
  });
}
int jets_scalar_destroy(jets_scalar s) {
  return guard([&] { JETS_CHECK(s, JETS_ERR_INVALID, "null scalar"); cudaFree(s->dev); delete s; });
}
int jets_scalar_set(jets_scalar s, double v) {
  return guard([&] {
    require_ready(); JETS_CHECK(s, JETS_ERR_INVALID, "null scalar");
    ctx().host_scratch[32] = v;
    CUDA_TRY(cudaMemcpyAsync(s->dev, &ctx().host_scratch[32], sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
    CUDA_TRY(cudaStreamSynchronize(ctx().stream));
  });
}
int jets_scalar_get(jets_scalar s, double* v) {
  return guard([&] {
    require_ready(); JETS_CHECK(s && v, JETS_ERR_INVALID, "null scalar");
    CUDA_TRY(cudaMemcpyAsync(&ctx().host_scratch[33], s->dev, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
    CUDA_TRY(cudaStreamSynchronize(ctx().stream));
    *v = ctx().host_scratch[33];
  });
}
int jets_dot_dev(jets_buf x, jets_buf y, jets_scalar out) {
  return guard([&] {
    require_ready(); check_pair(x, y); JETS_CHECK(out, JETS_ERR_INVALID, "null scalar");
    check_real(x->dtype, "jets_dot_dev");
    vec_reduce(x->dtype, 0, x->ptr(), y->ptr(), x->length(), 0, out->dev, ctx().stream);
  });
}
int jets_norm_dev(jets_buf x, double p, jets_scalar out) {
  return guard([&] { require_ready(); check_buf(x); JETS_CHECK(out, JETS_ERR_INVALID, "null scalar"); norm_to(x, p, out->dev); });
}
int jets_scalar_op(jets_scalar out, char op, jets_scalar a, jets_scalar b) {
  return guard([&] {
    require_ready(); JETS_CHECK(out && a, JETS_ERR_INVALID, "null scalar");
    scalar_op(out->dev, op, a->dev, b ? b->dev : nullptr, ctx().stream);
  });
}
int jets_axpby_dev(jets_buf out, jets_scalar sa, double ca, int af, jets_buf x, jets_scalar sb, double cb,
                   int bf, jets_buf y) {
  return guard([&] {
    require_ready(); check_pair(out, x);
    if (y) check_pair(out, y);
    check_real(out->dtype, "jets_axpby_dev");
    vec_axpby_dev(out->dtype, out->ptr(), out->length(), sa ? sa->dev : nullptr, ca, af, x->ptr(),
                  sb ? sb->dev : nullptr, cb, bf, y ? y->ptr() : nullptr, ctx().stream);
  });
}
int jets_axpby_pair_dev(jets_buf out1, jets_scalar s1a, double c1a, int f1a, jets_buf x1, jets_scalar s1b, double c1b, int f1b, jets_buf y1,
                        jets_buf out2, jets_scalar s2a, double c2a, int f2a, jets_buf x2, jets_scalar s2b, double c2b, int f2b, jets_buf y2) {
  return guard([&] {
    require_ready();
    JETS_CHECK(x1 && y1 && x2 && y2, JETS_ERR_INVALID, "null buffer handle");
    check_pair(out1, x1); check_pair(out1, y1); check_pair(out2, x2); check_pair(out2, y2); check_pair(out1, out2);
    check_real(out1->dtype, "jets_axpby_pair_dev");
    void* const out[2] = {out1->ptr(), out2->ptr()};
    const void* const x[2] = {x1->ptr(), x2->ptr()};
    const void* const y[2] = {y1->ptr(), y2->ptr()};
    const double* const sa[2] = {s1a ? s1a->dev : nullptr, s2a ? s2a->dev : nullptr};
    const double* const sb[2] = {s1b ? s1b->dev : nullptr, s2b ? s2b->dev : nullptr};
    const double ca[2] = {c1a, c2a}, cb[2] = {c1b, c2b};
    const int af[2] = {f1a, f2a}, bf[2] = {f1b, f2b};
    vec_axpby_pair_dev(out1->dtype, out1->length(), out, sa, ca, af, x, sb, cb, bf, y, ctx().stream);
  });
}
int jets_graph_begin(void) {
  return guard([&] {
    require_ready();
    CUDA_TRY(cudaStreamBeginCapture(ctx().stream, cudaStreamCaptureModeThreadLocal));
    ctx().capturing = true;
    ctx().capture_keep.clear();
  });
}
int jets_graph_end(void** exec_out) {
  return guard([&] {
    require_ready();
    cudaGraph_t g = nullptr;
    ctx().capturing = false;
    CUDA_TRY(cudaStreamEndCapture(ctx().stream, &g));
    cudaGraphExec_t e = nullptr;
    CUDA_TRY(cudaGraphInstantiate(&e, g, 0));
    cudaGraphDestroy(g);
    ctx().graph_keep[e] = std::move(ctx().capture_keep);
    ctx().capture_keep.clear();
    *exec_out = e;
  });
}
int jets_graph_launch(void* e) {
  return guard([&] { require_ready(); CUDA_TRY(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(e), ctx().stream)); });
}
int jets_graph_destroy(void* e) {
  return guard([&] {
    if (!e) return;
    cudaStreamSynchronize(ctx().stream);          // replays in flight still read the pinned plans
    cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(e));
    ctx().graph_keep.erase(e);
  });
}

// ------------------------------------------------------------------ leaves ---------------
int jets_op_diag(jets_buf w, jets_op* out) {
  return guard([&] {
    check_buf(w); JETS_CHECK(out, JETS_ERR_INVALID, "null out");
    jets_op a = new_op(K_DIAG, w->dtype);
    a->dom = a->rng = space1(w->length());
    w->refs++;
    a->w = w;
    *out = a;
  });
}
int jets_op_scale(jets_dtype dt, int64_t n, double s, jets_op* out) { return jets_op_scale_c(dt, n, s, 0.0, out); }
int jets_op_scale_c(jets_dtype dt, int64_t n, double re, double im, jets_op* out) {
  return guard([&] {
    check_dt(dt); JETS_CHECK(out && n >= 0, JETS_ERR_INVALID, "bad arguments");
    JETS_CHECK(im == 0.0 || is_cplx(dt), JETS_ERR_DTYPE, "complex scalar on a real space");
    jets_op a = new_op(K_SCALE, dt);
    a->dom = a->rng = space1(n);
    a->a = re; a->ai = im;
    *out = a;
  });
}
int jets_op_pointwise(jets_dtype dt, int64_t n, int fn, double p, jets_op* out) {
  return guard([&] {
    check_dt(dt); JETS_CHECK(out && n >= 0, JETS_ERR_INVALID, "bad arguments");
    JETS_CHECK(fn >= JETS_PW_SQUARE && fn <= JETS_PW_ATAN, JETS_ERR_UNSUPPORTED, "pointwise function %d is not in the registry", fn);
    JETS_CHECK(!is_cplx(dt) || fn == JETS_PW_SQUARE, JETS_ERR_UNSUPPORTED, "complex pointwise operators: only x^2 is in the registry");
    jets_op a = new_op(K_PW, dt);
    a->dom = a->rng = space1(n);
    a->fn = fn; a->p = p; a->linear = false;
    *out = a;
  });
}
int jets_op_stencil(jets_dtype dt, int64_t n, int kind, jets_op* out) {
  return guard([&] {
    check_dt(dt); JETS_CHECK(out && n >= 1, JETS_ERR_INVALID, "bad arguments");
    JETS_CHECK(kind == JETS_ST_FDIFF || kind == JETS_ST_LAP, JETS_ERR_UNSUPPORTED, "stencil kind %d is not in the registry", kind);
    jets_op a = new_op(K_STENCIL, dt);
    a->dom = a->rng = space1(n);
    a->fn = kind;
    *out = a;
  });
}
int jets_op_dense(jets_buf A, int64_t rows, int64_t cols, int64_t nrhs, jets_op* out) {
  return guard([&] {
    check_buf(A); JETS_CHECK(out && rows >= 1 && cols >= 1 && nrhs >= 1, JETS_ERR_INVALID, "bad arguments");
    JETS_CHECK(A->length() == rows * cols, JETS_ERR_SHAPE, "matrix buffer has %lld elements, expected %lld x %lld",
               (long long)A->length(), (long long)rows, (long long)cols);
    JETS_CHECK(rows < (1LL << 31) && cols < (1LL << 31), JETS_ERR_UNSUPPORTED, "matrix dimension too large");
    jets_op a = new_op(K_DENSE, A->dtype);
    a->dom = space1(cols * nrhs);
    a->rng = space1(rows * nrhs);
    a->rows = rows; a->cols = cols; a->nrhs = nrhs;
    A->refs++;
    a->w = A;
    *out = a;
  });
}
int jets_op_restrict(jets_dtype dt, int64_t n, int64_t nidx, const int64_t* idx0, jets_op* out) {
  return guard([&] {
    require_ready();
    check_dt(dt); JETS_CHECK(out && n >= 0 && nidx >= 0 && (idx0 || nidx == 0), JETS_ERR_INVALID, "bad arguments");
    std::vector<bool> seen((size_t)n, false);
    for (int64_t i = 0; i < nidx; ++i) {
      JETS_CHECK(idx0[i] >= 0 && idx0[i] < n, JETS_ERR_SHAPE, "restriction index %lld outside the domain [0,%lld)", (long long)idx0[i], (long long)n);
      JETS_CHECK(!seen[(size_t)idx0[i]], JETS_ERR_INVALID, "restriction index %lld appears twice (the adjoint would not be a plain scatter)", (long long)idx0[i]);
      seen[(size_t)idx0[i]] = true;
    }
    jets_op a = new_op(K_RESTRICT, dt);
    a->dom = space1(n);
    a->rng = space1(nidx);
    a->gidx64 = n >= (1LL << 31);
    const size_t bytes = (size_t)std::max<int64_t>(nidx, 1) * (a->gidx64 ? 8 : 4);
    void* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, bytes));
    a->gidx = std::shared_ptr<void>(d, [](void* p) { cudaFree(p); });
    if (nidx > 0) {
      if (a->gidx64) CUDA_TRY(cudaMemcpy(d, idx0, (size_t)nidx * 8, cudaMemcpyHostToDevice));
      else {
        std::vector<int32_t> h((size_t)nidx);
        for (int64_t i = 0; i < nidx; ++i) h[(size_t)i] = (int32_t)idx0[i];
        CUDA_TRY(cudaMemcpy(d, h.data(), (size_t)nidx * 4, cudaMemcpyHostToDevice));
      }
    }
    *out = a;
  });
}
int jets_op_zero(jets_dtype dt, int64_t ndom, int64_t nrng, jets_op* out) {
  return guard([&] {
    check_dt(dt); JETS_CHECK(out && ndom >= 0 && nrng >= 0, JETS_ERR_INVALID, "bad arguments");
    jets_op a = new_op(K_ZERO, dt);
    a->dom = space1(ndom);
    a->rng = space1(nrng);
    *out = a;
  });
}

// ------------------------------------------------------------------ combinators ----------
int jets_op_as_linear(jets_op a, jets_op* out) {
  return guard([&] { check_op(a); JETS_CHECK(out, JETS_ERR_INVALID, "null out"); *out = as_linear(a); });
}
int jets_op_adjoint(jets_op a, jets_op* out) {
  return guard([&] { check_op(a); JETS_CHECK(out, JETS_ERR_INVALID, "null out"); *out = adjoint_of(a); });
}

// jops_comp (src/Jets.jl:542-550): the flattened operand list of `a`, each handle retained.
static void comp_operands(jets_op a, std::vector<jets_op>& out) {
  if (a->kind == K_COMPOSE) {
    for (jets_op k : a->kids) out.push_back(retain(k));
  } else if (a->kind == K_LNVIEW && a->kids[0]->kind == K_COMPOSE) {
    for (jets_op k : a->kids[0]->kids) out.push_back(as_linear(k));
  } else if (a->kind == K_ADJ && strip(a)->kind == K_COMPOSE && a->kids[0]->kind != K_ADJ) {
    jets_op c = strip(a);
    for (auto it = c->kids.rbegin(); it != c->kids.rend(); ++it) {
      jets_op l = as_linear(*it);
      out.push_back(adjoint_of(l));
      op_release(l);
    }
  } else {
    out.push_back(retain(a));
  }
}
int jets_op_compose(int32_t n, const jets_op* ops, jets_op* out) {
  return guard([&] {
    JETS_CHECK(n >= 1 && ops && out, JETS_ERR_INVALID, "bad arguments");
    std::vector<jets_op> flat;
    bool lin = true;
    for (int i = 0; i < n; ++i) {
      check_op(ops[i]);
      lin = lin && ops[i]->linear;
      comp_operands(ops[i], flat);
    }
    jets_op a = new_op(K_COMPOSE, ops[0]->dtype);
    a->kids = flat;
    try {
      for (size_t i = 0; i < flat.size(); ++i) {
        JETS_CHECK(flat[i]->dtype == a->dtype, JETS_ERR_DTYPE, "composition of operators with different eltypes");
        if (i + 1 < flat.size())
          JETS_CHECK(flat[i]->dom.total() == flat[i + 1]->rng.total(), JETS_ERR_SHAPE,
                     "composition: domain of operand %d (%lld) does not match range of operand %d (%lld)", (int)i,
                     (long long)flat[i]->dom.total(), (int)i + 1, (long long)flat[i + 1]->rng.total());
      }
    } catch (...) { op_release(a); throw; }
    a->dom = flat.back()->dom;   // src/Jets.jl:522
    a->rng = flat.front()->rng;
    a->linear = lin;
    *out = a;
  });
}

static void sum_operands(jets_op a, int sign, std::vector<jets_op>& ops, std::vector<int>& sg) {
  if (a->kind == K_SUM) {
    for (size_t k = 0; k < a->kids.size(); ++k) { ops.push_back(retain(a->kids[k])); sg.push_back(a->sgn[k] * sign); }
  } else if (a->kind == K_LNVIEW && a->kids[0]->kind == K_SUM) {
    jets_op s = a->kids[0];
    for (size_t k = 0; k < s->kids.size(); ++k) { ops.push_back(as_linear(s->kids[k])); sg.push_back(s->sgn[k] * sign); }
  } else if (a->kind == K_ADJ && strip(a)->kind == K_SUM && a->kids[0]->kind != K_ADJ) {
    jets_op s = strip(a);
    for (size_t k = 0; k < s->kids.size(); ++k) {
      jets_op l = as_linear(s->kids[k]);
      ops.push_back(adjoint_of(l));
      op_release(l);
      sg.push_back(s->sgn[k] * sign);
    }
  } else {
    ops.push_back(retain(a));
    sg.push_back(sign);
  }
}
int jets_op_sum(int32_t n, const jets_op* ops, const int32_t* sgn, jets_op* out) {
  return guard([&] {
    JETS_CHECK(n >= 1 && ops && sgn && out, JETS_ERR_INVALID, "bad arguments");
    std::vector<jets_op> flat;
    std::vector<int> sg;
    bool lin = true;
    for (int i = 0; i < n; ++i) {
      check_op(ops[i]);
      JETS_CHECK(sgn[i] == 1 || sgn[i] == -1, JETS_ERR_INVALID, "signs must be +1/-1");
      lin = lin && ops[i]->linear;
      sum_operands(ops[i], sgn[i], flat, sg);
    }
    jets_op a = new_op(K_SUM, ops[0]->dtype);
    a->kids = flat;
    a->sgn = sg;
    try {
      for (jets_op k : flat) {
        JETS_CHECK(k->dtype == a->dtype, JETS_ERR_DTYPE, "sum of operators with different eltypes");
        JETS_CHECK(k->dom.total() == flat[0]->dom.total() && k->rng.total() == flat[0]->rng.total(), JETS_ERR_SHAPE,
                   "sum of operators with inconsistent domains/ranges");
      }
    } catch (...) { op_release(a); throw; }
    a->dom = flat[0]->dom;  // src/Jets.jl:628
    a->rng = flat[0]->rng;
    a->linear = lin;
    *out = a;
  });
}

int jets_op_block(int32_t R, int32_t C, const jets_op* ops, int dadom, jets_op* out) {
  return guard([&] {
    JETS_CHECK(R >= 1 && C >= 1 && ops && out, JETS_ERR_INVALID, "bad arguments");
    jets_op a = new_op(K_BLOCK, ops[0] ? ops[0]->dtype : JETS_F64);
    a->R = R; a->C = C;
    bool lin = true;
    try {
      for (int i = 0; i < R * C; ++i) {
        check_op(ops[i]);
        a->kids.push_back(retain(ops[i]));
        lin = lin && ops[i]->linear;
        JETS_CHECK(ops[i]->dtype == a->dtype, JETS_ERR_DTYPE, "block operator with mixed eltypes");
        JETS_CHECK(ops[i]->dom.len.size() == 1 && ops[i]->rng.len.size() == 1, JETS_ERR_UNSUPPORTED,
                   "block operators as blocks of a block operator are rejected: the reference's own forward is wrong for them "
                   "(JetBlock_df! accumulates into a temporary shared by the block columns, src/Jets.jl:1012-1024)");
      }
      // spaces from the first block row / column (src/Jets.jl:927-928); every block must agree
      for (int c = 0; c < C; ++c) a->dom.len.push_back(ops[(size_t)c * R]->dom.total());
      for (int r = 0; r < R; ++r) a->rng.len.push_back(ops[r]->rng.total());
      for (int c = 0; c < C; ++c)
        for (int r = 0; r < R; ++r) {
          jets_op k = ops[r + (size_t)c * R];
          JETS_CHECK(k->dom.total() == a->dom.len[c] && k->rng.total() == a->rng.len[r], JETS_ERR_SHAPE,
                     "block (%d,%d) maps %lld -> %lld but its column/row spaces are %lld -> %lld", r, c,
                     (long long)k->dom.total(), (long long)k->rng.total(), (long long)a->dom.len[c], (long long)a->rng.len[r]);
        }
    } catch (...) { op_release(a); throw; }
    a->dom.is_block = (C > 1) || dadom;
    a->rng.is_block = true;
    a->linear = lin;
    *out = a;
  });
}

int jets_op_scalar_mul(double s, jets_op A, jets_op* out) { return jets_op_scalar_mul_c(s, 0.0, A, out); }
int jets_op_scalar_mul_c(double s, double s_im, jets_op A, jets_op* out) {
  return guard([&] {
    check_op(A); JETS_CHECK(out, JETS_ERR_INVALID, "null out");
    jets_op sc = nullptr;
    { const int rc = jets_op_scale_c((jets_dtype)A->dtype, A->rng.total(), s, s_im, &sc); if (rc != JETS_OK) throw Fail{rc}; }
    jets_op pair[2] = {sc, A};
    jets_op res = nullptr;
    const int rc = jets_op_compose(2, pair, &res);
    op_release(sc);
    if (rc != JETS_OK) throw Fail{rc};
    *out = res;
  });
}

// ------------------------------------------------------------------ queries --------------
int jets_op_retain(jets_op a) { return guard([&] { check_op(a); a->refs++; }); }
int jets_op_destroy(jets_op a) { return guard([&] { check_op(a); op_release(a); }); }
int jets_op_is_linear(jets_op a) { return a ? (a->linear ? 1 : 0) : -1; }
int jets_op_is_zero(jets_op a) { return a ? (strip(a)->kind == K_ZERO ? 1 : 0) : -1; }
int jets_op_is_block(jets_op a) { return a ? (strip(a)->kind == K_BLOCK ? 1 : 0) : -1; }
int jets_op_dtype(jets_op a) { return a ? a->dtype : -1; }
int32_t jets_op_nblocks(jets_op a, int which) {
  if (!a) return -1;
  return (int32_t)(which == 1 ? a->rng.len.size() : a->dom.len.size());
}
int jets_op_block_len(jets_op a, int which, int32_t b, int64_t* len) {
  return guard([&] {
    check_op(a);
    const Space& s = which == 1 ? a->rng : a->dom;
    JETS_CHECK(b >= 0 && b < (int)s.len.size() && len, JETS_ERR_INVALID, "block %d out of range", b);
    *len = s.len[b];
  });
}

static jets_op getblock_impl(jets_op a, int i, int j) {  // retained result
  if (a->kind == K_ADJ) {  // getblock(A', i, j) = getblock(A, j, i)'  (src/Jets.jl:1088)
    jets_op b = getblock_impl(a->kids[0], j, i);
    jets_op l = as_linear(b);
    jets_op r = adjoint_of(l);
    op_release(b); op_release(l);
    return r;
  }
  if (a->kind == K_LNVIEW) {  // :1086
    jets_op b = getblock_impl(a->kids[0], i, j);
    jets_op l = as_linear(b);
    op_release(b);
    return l;
  }
  if (a->kind == K_BLOCK) {
    JETS_CHECK(i >= 0 && i < a->R && j >= 0 && j < a->C, JETS_ERR_INVALID, "block (%d,%d) out of range %dx%d", i, j, a->R, a->C);
    return retain(a->kids[i + (size_t)j * a->R]);
  }
  if (a->kind == K_COMPOSE) {  // :1100-1110
    std::vector<jets_op> parts;
    for (jets_op k : a->kids) parts.push_back(strip(k)->kind == K_BLOCK ? getblock_impl(k, i, j) : retain(k));
    jets_op res = nullptr;
    const int rc = jets_op_compose((int)parts.size(), parts.data(), &res);
    for (jets_op p : parts) op_release(p);
    if (rc != JETS_OK) throw Fail{rc};
    return res;
  }
  JETS_FAIL(JETS_ERR_INVALID, "getblock: not a block operator");
}
int jets_op_getblock(jets_op a, int32_t i, int32_t j, jets_op* out) {
  return guard([&] { check_op(a); JETS_CHECK(out, JETS_ERR_INVALID, "null out"); *out = getblock_impl(a, i, j); });
}

// ------------------------------------------------------------------ linearization --------
int jets_op_set_point(jets_op a, jets_buf mo) {
  return guard([&] {
    require_ready(); check_op(a); check_buf(mo);
    JETS_CHECK(mo->length() == a->dom.total() || a->kind == K_ADJ, JETS_ERR_SHAPE, "point has %lld elements, domain has %lld",
               (long long)mo->length(), (long long)a->dom.total());
    set_point_impl(a, mo);
    g_epoch++;
  });
}
int jets_op_jacobian(jets_op a, jets_buf mo, jets_op* out) {
  return guard([&] {
    require_ready(); check_op(a); check_buf(mo); JETS_CHECK(out, JETS_ERR_INVALID, "null out");
    if (a->linear) {  // jacobian of a linear operator is the operator (src/Jets.jl:366): share it
      *out = retain(a);
      return;
    }
    JETS_CHECK(mo->length() == a->dom.total(), JETS_ERR_SHAPE, "point has %lld elements, domain has %lld",
               (long long)mo->length(), (long long)a->dom.total());
    jets_op c = clone_tree(a);
    jets_buf snap = nullptr;
    try {
      std::vector<int64_t> len(mo->nblocks());
      for (int i = 0; i < mo->nblocks(); ++i) len[i] = mo->blk_off[i + 1] - mo->blk_off[i];
      snap = alloc_buf(mo->dtype, mo->nblocks(), len.data());  // copy(mo), src/Jets.jl:374
      CUDA_TRY(cudaMemcpyAsync(snap->ptr(), mo->ptr(), (size_t)mo->length() * dsize(mo->dtype), cudaMemcpyDeviceToDevice, ctx().stream));
      set_point_impl(c, snap);
      g_epoch++;
      buf_release(snap);
    } catch (...) { op_release(c); if (snap) buf_release(snap); throw; }
    jets_op l = as_linear(c);
    op_release(c);
    *out = l;
  });
}

int jets_op_clone(jets_op a, jets_op* out) {
  return guard([&] {
    check_op(a); JETS_CHECK(out, JETS_ERR_INVALID, "null out");
    *out = clone_tree(a);   // new nodes (own plan cache, own linearization point slots), shared state buffers
  });
}

// ------------------------------------------------------------------ apply ----------------
int jets_apply(jets_op a, int mode, jets_buf out, jets_buf in, int accumulate) {
  return guard([&] { apply_impl(a, mode, out, in, accumulate); });
}
int jets_apply_axpby(jets_op a, int mode, jets_buf out, jets_buf in, jets_scalar sa, double ca, int a_flags,
                     jets_scalar so, double co, int o_flags) {
  return guard([&] {
    ApplyCoef c;
    c.a_ptr = sa ? sa->dev : nullptr; c.a_const = ca; c.a_flags = a_flags;
    c.o_ptr = so ? so->dev : nullptr; c.o_const = co; c.o_flags = o_flags;
    apply_impl(a, mode, out, in, 0, &c);
  });
}
int jets_apply_axpby_norm(jets_op a, int mode, jets_buf out, jets_buf in, jets_scalar sa, double ca, int a_flags,
                          jets_scalar so, double co, int o_flags, jets_scalar norm2_out) {
  return guard([&] {
    JETS_CHECK(norm2_out, JETS_ERR_INVALID, "null scalar");
    ApplyCoef c;
    c.a_ptr = sa ? sa->dev : nullptr; c.a_const = ca; c.a_flags = a_flags;
    c.o_ptr = so ? so->dev : nullptr; c.o_const = co; c.o_flags = o_flags;
    apply_impl(a, mode, out, in, 0, &c, norm2_out->dev);
  });
}
int jets_scalar_prog(int32_t n, const jets_scalar* out, const char* op, const jets_scalar* a, const jets_scalar* b) {
  return guard([&] {
    require_ready();
    JETS_CHECK(n >= 0 && out && op && a && b, JETS_ERR_INVALID, "null argument");
    for (int32_t i0 = 0; i0 < n; i0 += kMaxScalarProg) {
      ScalarProg p{};
      p.n = std::min<int32_t>(kMaxScalarProg, n - i0);
      for (int i = 0; i < p.n; ++i) {
        JETS_CHECK(out[i0 + i] && a[i0 + i], JETS_ERR_INVALID, "null scalar in program step %d", i0 + i);
        p.out[i] = out[i0 + i]->dev;
        p.a[i] = a[i0 + i]->dev;
        p.b[i] = b[i0 + i] ? b[i0 + i]->dev : nullptr;
        p.op[i] = op[i0 + i];
      }
      scalar_prog(p, ctx().stream);
    }
  });
}
int jets_op_plan_info(jets_op a, int mode, int32_t* engines, int32_t* nlaunches) {
  return guard([&] {
    check_op(a);
    int e = 0, n = 0;
    for (auto& kv : a->plans)
      if ((kv.first & 3) == mode) { e |= kv.second->engines; n = (int)kv.second->steps.size(); }
    if (engines) *engines = e;
    if (nlaunches) *nlaunches = n;
  });
}

}  // extern "C"
