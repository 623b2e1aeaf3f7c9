// Internal declarations shared by the translation units of libjets_b200.so.
// Nothing here is part of the ABI (include/jets_b200.h is).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jets_b200.h"

namespace jets {

// ------------------------------------------------------------------ errors ---------------
void set_error(const char* fmt, ...);
struct Fail {
  int code;
};
#define JETS_FAIL(code, ...)        \
  do {                              \
    ::jets::set_error(__VA_ARGS__); \
    throw ::jets::Fail{code};       \
  } while (0)
#define JETS_CHECK(cond, code, ...)              \
  do {                                           \
    if (!(cond)) JETS_FAIL(code, __VA_ARGS__);   \
  } while (0)
#define CUDA_TRY(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
      JETS_FAIL(JETS_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__),   \
                __FILE__, __LINE__, #expr);                                              \
  } while (0)

// Wraps an ABI body: converts Fail / std exceptions into status codes.
template <class F>
int guard(F&& f) {
  try {
    f();
    return JETS_OK;
  } catch (const Fail& e) {
    return e.code;
  } catch (const std::exception& e) {
    set_error("internal error: %s", e.what());
    return JETS_ERR_INVALID;
  }
}

// ------------------------------------------------------------------ context --------------
struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // fork/join onto auxiliary streams (communication overlapped with compute, dist.py)
  cudaStream_t main_stream = nullptr;   // saved while the context stream is an auxiliary one
  cudaStream_t aux[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  bool on_aux = false;
  int64_t launches = 0;
  int fused_engine = 0;  // 0 auto, 1 TMA, 2 LDG
  int fast_variant = -1; // JETS_B200_FAST_VARIANT (-1 = chosen per plan): consumer shape of the fast TMA kernel
  int no_fast = 0;       // JETS_B200_NO_FAST=1: always use the interpreter kernel
  int no_bundle = 0;     // JETS_B200_NO_BUNDLE=1: fast kernel without the shared-memory input-tile cache
  int bundle_nx = 0;     // JETS_B200_BUNDLE_NX / _NS / _BMAX: override the planner's ring sizes (tuning)
  int bundle_ns = 0;
  int bundle_bmax = 0;
  int no_pdl = 0;
  int vec_pdl = 0;        // JETS_B200_VEC_PDL=1: the vector kernels (reductions, axpby, scalar programs) launch with PDL too (A/B: measured slower)
  // JETS_B200_TRACE=1: every bundle launch writes a per-CTA timeline (globaltimer) into a ring of launch records
  static constexpr int kTraceLaunches = 64, kTraceCtas = 160;
  unsigned long long* trace_buf = nullptr;
  int64_t trace_count = 0;
  int no_first_static = 0;   // JETS_B200_NO_FIRST_STATIC=1: the first claim of a CTA goes through the atomic counter too (A/B)
  int64_t tail_min_units = 0; // JETS_B200_TAIL_MIN_UNITS: split the tail of launches with at least this many full-height units (tests: 1)
  int tail_min_rows = 16;    // JETS_B200_TAIL_MIN_ROWS: bundles with at least this many rows are cut into sub-bundles at the tail
  int tail_div = 8;          // JETS_B200_TAIL_DIV: sub-bundles of 1/div of the rows ...
  int tail_sub_min = 8;      // JETS_B200_TAIL_SUB_MIN: ... but at least this many
  int64_t tile_elems = 0;    // JETS_B200_TILE_ELEMS: force the tile length of the bundle kernel (tuning; rounded to 128 bytes, capped by the buffer)
  int no_tail_split = 0;     // JETS_B200_NO_TAIL_SPLIT=1: no fine-grained sub-bundles over the last tile positions (A/B)
  int group_streams = 0;     // JETS_B200_GROUP_STREAMS=n: at most n state streams per term group (tuning)
  int no_pre_state = 0;      // JETS_B200_NO_PRE_STATE=1: never fetch operator state before griddepcontrol.wait (A/B)
  uintptr_t pdl_out_lo = 0, pdl_out_hi = 0;   // what the last launch that lets its dependents start early (bundle / PDL vector kernel) writes
  int dist_early_ctas = 16;  // JETS_B200_DIST_EARLY_CTAS: CTAs that take the peer-store units of a distributed apply (0: all)
  int grid_limit = 0;    // JETS_B200_GRID=n: launch the fused kernels with at most n CTAs (leaves SMs to concurrent kernels)
  int static_sched = 0;  // JETS_B200_STATIC_SCHED=1: deal units round-robin instead of claiming them dynamically        // JETS_B200_NO_PDL=1: launch without programmatic stream serialization
  double* host_scratch = nullptr;  // pinned, 64 doubles
  double* dev_scratch = nullptr;   // device partials for reductions
  size_t dev_scratch_elems = 0;
  bool capturing = false;
  // Plans launched while a CUDA graph is being captured: the graph holds raw pointers into their tables and
  // temporaries, so the executable graph keeps them alive (a later point! may replace the operator's cached plan).
  std::vector<std::shared_ptr<void>> capture_keep;
  std::map<void*, std::vector<std::shared_ptr<void>>> graph_keep;
};
Context& ctx();
void require_ready();
inline size_t dsize(int dt) { return dt == JETS_F32 ? 4 : dt == JETS_C128 ? 16 : 8; }
inline bool is_cplx(int dt) { return dt == JETS_C64 || dt == JETS_C128; }
inline int real_of(int dt) { return dt == JETS_C64 ? JETS_F32 : dt == JETS_C128 ? JETS_F64 : dt; }

// ------------------------------------------------------------------ storage --------------
constexpr size_t kGuardBytes = 256;  // readable slack before/after every owned allocation

struct Storage {
  void* alloc = nullptr;  // cudaMalloc'ed base (owned) or nullptr (wrapped)
  char* data = nullptr;   // first logical byte
  size_t bytes = 0;       // logical bytes
  bool guarded = false;   // kGuardBytes readable on both sides
  ~Storage();
};

}  // namespace jets

// Handle types live in the global namespace (they are the ABI's opaque structs).
struct jets_buf_s {
  int refs = 1;
  int dtype = JETS_F64;
  std::shared_ptr<jets::Storage> st;
  int64_t off = 0;                 // element offset of this view inside st
  std::vector<int64_t> blk_off;    // nblocks+1 cumulative element offsets (0-based), relative
  int64_t length() const { return blk_off.back(); }
  int32_t nblocks() const { return (int32_t)blk_off.size() - 1; }
  char* ptr() const { return st->data + (size_t)off * jets::dsize(dtype); }
  char* block_ptr(int b) const { return ptr() + (size_t)blk_off[b] * jets::dsize(dtype); }
  bool guarded() const { return st->guarded; }
};

struct jets_scalar_s {
  double* dev = nullptr;
};

namespace jets {

// A JetSpace (one block, is_block=false) or a JetBSpace (is_block=true).
struct Space {
  std::vector<int64_t> len;  // per-block lengths
  bool is_block = false;
  int64_t total() const {
    int64_t s = 0;
    for (auto l : len) s += l;
    return s;
  }
  bool same_layout(const Space& o) const { return len == o.len; }
};

enum Kind : int {
  K_DIAG, K_SCALE, K_PW, K_STENCIL, K_DENSE, K_ZERO, K_RESTRICT,  // leaves
  K_LNVIEW, K_ADJ,                                    // wrappers
  K_COMPOSE, K_SUM, K_BLOCK                           // combinators
};

struct Plan;

}  // namespace jets

struct jets_op_s {
  int refs = 1;
  jets::Kind kind;
  int dtype = JETS_F64;
  jets::Space dom, rng;
  bool linear = true;  // JopLn-like (df! == f!) or an explicit linear view
  // leaf state
  jets_buf w = nullptr;      // diagonal / dense matrix (retained)
  double a = 0, p = 0;       // scale constant / pointwise parameter
  double ai = 0;             // imaginary part of the scale constant (complex spaces)
  int fn = 0;                // pointwise fn or stencil kind
  int64_t rows = 0, cols = 0, nrhs = 1;
  jets_buf mo = nullptr;     // linearization point of a pointwise leaf (retained, by reference)
  std::shared_ptr<void> gidx;  // K_RESTRICT: device index table (int32 when the domain allows, else int64), shared by clones
  bool gidx64 = false;
  // children
  std::vector<jets_op> kids;  // retained
  std::vector<int> sgn;       // K_SUM
  int R = 0, C = 0;           // K_BLOCK (kids column-major)
  // plan cache, keyed by mode | accumulate<<2 | engine<<3 ; invalidated when `version` bumps
  uint64_t version = 0;
  std::map<int, std::shared_ptr<jets::Plan>> plans;
  void* axpby_tmp = nullptr;      // staging of jets_apply_axpby for plans that are not one fused launch
  size_t axpby_tmp_bytes = 0;
  ~jets_op_s();
};

namespace jets {

// ------------------------------------------------------------------ fused engine tables --
// One "stage" of an elementwise/stencil chain, interpreted per 128-bit vector.
enum StageOp : int {
  S_SCALE = 0,   // v *= c0
  S_DIAG = 1,    // v *= stream[p]
  S_PW_F = 2,    // v = phi(v)
  S_PW_J = 3,    // v = phi'(stream[p]) * v
  S_FDIFF = 4,   // v[p] = p+1<n ? v[p+1]-v[p] : 0
  S_BDIFF = 5,   // adjoint of FDIFF: v[p] = (p>=1 ? v[p-1] : 0) - (p+1<n ? v[p] : 0)
  S_LAP = 6,     // v[p] = ((p>=1?v[p-1]:0) - 2 v[p]) + (p+1<n?v[p+1]:0)
  S_NEG = 7,     // v = -v
  S_ZERO = 8,    // v = 0
  S_CSCALE = 9,  // complex spaces: v *= (c0 + i*c0 of the S_CIMAG stage that follows)
  S_CIMAG = 10   // carries the imaginary part of the preceding S_CSCALE; no operation
};

struct FStage {      // 32 bytes
  int32_t op;
  int32_t fn;
  const void* ptr;   // operand stream, position 0 of the block (absolute), or null
  double c0;
  double c1;
};
struct FTerm {       // 32 bytes
  int64_t in_off;    // element offset of the input block relative to the `in` base
  const void* in_abs;  // absolute input pointer (overrides in_off when non-null)
  int32_t stage_begin, stage_end;
  int32_t sign;      // +1 / -1
  int32_t nstreams;  // 1 (input) + stages with ptr
};
struct FRow {        // 40 bytes
  int64_t out_off;   // element offset of the output block relative to the `out` base
  int64_t len;
  int32_t term_begin, term_end;
  int32_t init;      // 0 zero-init, 1 accumulate onto existing out, 2 leave untouched
  int32_t ntiles;
  int32_t group_begin, group_end;  // term groups (one TMA slot each)
};
// Compact, kernel-facing description of a term GROUP: consecutive terms of one output row whose
// operand streams share one TMA slot.  Lives in global memory (read-only, L1/L2 resident): the
// producer reads ptr/nstreams, the consumers read terms/stages.
constexpr int kMaxStages = 6;     // per term (longer chains are split by the planner)
constexpr int kMaxStreams = 4;    // operand streams per slot (= per term group)
constexpr int kGroupTerms = 4;    // terms of one output tile that share a slot
constexpr int kGroupStages = 12;  // stage pool of a group
struct CStage {      // 16 bytes
  uint8_t op, fn, has_stream, pad0;
  uint32_t pad1;
  double c0;
};
// Chains the consumer evaluates with straight-line, compile-time-specialised code instead of the
// stage interpreter (same IEEE operations in the same order, so results are bit-identical).
enum Pattern : int {
  PAT_GENERIC = 0,
  PAT_COPY,           // no stage (identity)
  PAT_DIAG,           // w .* x                      (JopBlock of diagonal JopLn, config 1)
  PAT_FDIFF, PAT_BDIFF, PAT_LAP,   // bare stencils
  PAT_SCALE,          // c .* x
  PAT_J2_FDIFF_DIAG,  // w .* S(2 mo .* x)           (config 2 forward: D ∘ S ∘ J)
  PAT_DIAG_BDIFF_J2,  // 2 mo .* S'(w .* x)          (config 2 adjoint)
  PAT_LAP_SCALE, PAT_FDIFF_SCALE, PAT_BDIFF_SCALE,   // c .* S(x)   (config 4: B - c*S)
  PAT_J2,             // 2 mo .* x                   (Jacobian of x^2)
  PAT_SQUARE,         // x .* x
  PAT_SCALE_LAP, PAT_SCALE_FDIFF, PAT_SCALE_BDIFF,   // S(c .* x)   (adjoint of c*S: config 4's A')
  PAT_LAP_DIAG, PAT_FDIFF_DIAG, PAT_BDIFF_DIAG,      // w .* S(x)   (weighted derivative blocks)
  PAT_DIAG_LAP, PAT_DIAG_FDIFF, PAT_DIAG_BDIFF       // S(w .* x)   (their adjoints)
};
struct GTerm {       // 8 bytes
  uint8_t stage0, nstages;   // into the group's stage pool
  uint8_t stream0;           // first operand stream of the term inside the slot
  uint8_t pattern;           // Pattern
  int8_t sign;
  uint8_t pad[3];
};
struct GroupRec {    // 272 bytes
  int64_t ptr[kMaxStreams];  // absolute address, or byte offset from the apply's `in` base (rel_mask bit)
  int32_t nstreams, nterms, rel_mask, pad;
  GTerm terms[kGroupTerms];
  CStage stages[kGroupStages];
};
// ---- bundle engine (kernels_fused_bundle.cu): rows that share input blocks are walked by ONE CTA
// at the same tile position, and every input tile lives in a shared-memory ring ("x ring") from
// its first to its last use, so a block-tridiagonal row costs 2 tile loads instead of 4.
enum : int { XF_LOAD = 1, XF_RELEASE = 2 };
// BGroupRec::flags: bits 4-5 select the output base (0 = the apply's `out`, k = GateLaunch::out_alt[k-1]).
// (BG_SIG_SHIFT: where the kernel's slot metadata carries a bundle's signal mask.)
enum : int { BG_ROW_FIRST = 1, BG_ROW_LAST = 2, BG_ACC = 4, BG_OUT_ALT_SHIFT = 4, BG_SIG_SHIFT = 8 };
// BGroupRec::xrel_mask: bit t = input of term t is an offset from a base pointer; bits 8+2t..9+2t select
// that base (0 = the apply's `in`, k = GateLaunch::in_alt[k-1]).
constexpr int kXAltShift = 8;
constexpr int kGateFlags = 4;     // cross-rank flag words a launch may wait on / signals it may raise
constexpr int kGateFlagStride = 32;  // 32-bit words between two flag words (one 128-byte line each)
struct BTerm {       // 8 bytes
  uint8_t stage0, nstages;   // into the group's stage pool
  uint8_t sstream0;          // first STATE stream of the term inside the state slot
  uint8_t pattern;           // Pattern
  int8_t sign;
  uint8_t xflags;            // bit0 XF_LOAD (first use of the input tile), bit1 XF_RELEASE (last use),
                             // bit2 = (xrel / NX) & 1, bits 4-7 = xrel % NX  (filled in once NX is final)
  uint16_t xrel;             // allocation index of the input tile inside the unit (x ring slot = (base+xrel) % NX)
};
struct BGroupRec {   // 320 bytes
  int64_t xptr[kGroupTerms];   // input block of each term: byte offset from the apply's `in` (xrel_mask bit) or absolute
  int64_t sptr[kMaxStreams];   // state streams (absolute addresses)
  int32_t nsstreams, nterms, xrel_mask, flags;   // flags: BG_*   (16B aligned: fetched with one 128-bit load)
  BTerm terms[kGroupTerms];    // 16B aligned
  CStage stages[kGroupStages]; // 16B aligned
  int64_t out_off;             // element offset of the output row relative to the apply's `out`
  int32_t row_in_bundle, pad;  // which row of its bundle the group belongs to (norm partial index)
};
static_assert(offsetof(BGroupRec, nsstreams) % 16 == 0 && offsetof(BGroupRec, terms) % 16 == 0 &&
              offsetof(BGroupRec, stages) % 16 == 0, "BGroupRec alignment");
static_assert(sizeof(BGroupRec) == 320, "BGroupRec layout");
struct BundleRec {   // 64 bytes: consecutive output rows of equal length walked by one CTA per tile position
  int64_t unit_begin;          // first (bundle, position) unit of this bundle in the launch-wide enumeration
  int64_t len;                 // row length (elements)
  int32_t group_begin, ngroups;
  int32_t nx;                  // x-ring allocations per unit
  int32_t gate;                // bits 0-3: flag words the producer waits for before the unit's first load;
                               // bits 4-7: signals a finished unit of this bundle counts towards
  int32_t claim_begin, chunk;  // dynamic scheduling: first claim of this bundle, units per claim
  int64_t pbase;               // first (unit, row) tile of this bundle in the launch-wide enumeration
  int32_t nrows, nclaims;      // rows of the bundle; dynamic claims it is cut into (ceil(npos / chunk))
  int32_t pos0, npos;          // the tile positions this record covers: [pos0, pos0 + npos) of ceil(len / tile)
};

struct FSeg {        // schedule segment: positions [pos_begin, ...) with `nactive` rows active
  int64_t tile_begin;
  int64_t pos_begin;
  int32_t nactive;
  int32_t pad;
};

struct FusedTables {
  std::vector<FStage> stages;
  std::vector<FTerm> terms;
  std::vector<FRow> rows;
  std::vector<GroupRec> groups;
  int hl = 0, hr = 0;     // halo (elements) the chains need on the left / right
  int max_streams = 1;
  bool tma_ok = false;    // all streams 16B aligned + guarded
};

struct DevFused {   // device copy + launch geometry
  FStage* stages = nullptr;
  FTerm* terms = nullptr;
  FRow* rows = nullptr;
  GroupRec* groups = nullptr;
  FSeg* segs = nullptr;
  int32_t* order = nullptr;  // rows sorted by ntiles (desc)
  int32_t nrows = 0, nsegs = 0;
  int64_t ntiles = 0;   // real tiles
  int64_t nitems = 0;   // (row, chunk-of-tiles) work items
  int hl = 0, hr = 0, slot_streams = 1;
  int S = 1;            // tiles per (super-chunk, row) item
  int tile_elems = 0;
  bool use_tma = false;
  bool heavy = false;   // chains use transcendental pointwise functions
  bool fast = false;    // every chain has a straight-line fast path -> jets_fused_fast_kernel
  int variant = 0;      // fast-kernel shape: 0 = 16 warps x 1 vec, 1 = 8 x 2, 2 = 16 x 2 (16 KB tiles)
  // bundle engine
  bool bundle = false;
  BGroupRec* bgroups = nullptr;
  BundleRec* bundles = nullptr;
  int32_t nbundles = 0, NX = 0, NS = 0, G = 1, sstreams = 0;
  int64_t nunits = 0;
  size_t table_bytes = 0;
  int32_t* sched = nullptr;   // dynamic unit scheduler counters (in the plan blob) or null
  int32_t chunk = 1;
  bool covers_out = false;    // the launch writes every element of the apply's `out`
  void* blob = nullptr;
  // cross-rank gating (distributed banded apply, dist.cu): units per signal, the signals this launch must
  // raise even when no unit feeds them, and the completion counters (in the plan blob)
  // byte range the launch writes relative to the apply's `out`, and the absolute range of its operator-state
  // streams: a launch may fetch state before griddepcontrol.wait unless the previous launch wrote into that range
  int64_t out_lo = 0, out_hi = 0;
  uintptr_t state_lo = 0, state_hi = 0;
  int64_t nclaims = 0;        // dynamic claims (bundle-major)
  int64_t early_claims = 0;   // leading claims whose units store to peer memory (taken by a few CTAs only)
  int64_t nrowtiles = 0;      // (unit, row) tiles of the launch: the store epilogue can leave one sum-of-squares partial
                              // per tile and consumer warp for a fixed-order norm (jets_apply_axpby_norm)
  int consumer_warps = 16;
  int32_t sig_total[kGateFlags] = {0, 0, 0, 0};
  int32_t sig_owned = 0;
  int32_t* sig_done = nullptr;
  bool gated = false;
};

// Dense block table entry: out[out_off + i] (+)= sum_j A[i + j*lda] * in[in_off + j]  (trans=0)
//                          out[out_off + j] (+)= sum_i A[i + j*lda] * in[in_off + i]  (trans=1)
struct DBlock {
  const void* A;
  int64_t in_off, out_off;
  int32_t rows, cols;   // of the stored matrix
  int32_t lda;
  int32_t trans;
  int32_t nrhs;
  int32_t pad;
};

enum StepKind : int { ST_FUSED, ST_GEMV, ST_FILL0, ST_GEMM_TC, ST_GATHER };
enum AccMode : int { ACC_SET = 0, ACC_ADD = 1, ACC_SUB = 2 };

struct Ref {           // where a step reads / writes: 0 = apply's `in`, 1 = apply's `out`, >=2 tmp
  int which = 0;
  int64_t off = 0;     // element offset inside that buffer
};

struct Step {
  StepKind kind;
  Ref src, dst;
  DevFused fused;              // ST_FUSED
  std::vector<DBlock> dblocks; // ST_GEMV (host copy)
  DBlock* d_dblocks = nullptr;
  int32_t* d_row_ptr = nullptr;  // output-row grouping for GEMV
  int32_t n_out_rows = 0;
  int64_t gemv_tiles = 0;
  int32_t gemv_ksplit = 1;         // N orientation with few row tiles: CTAs per tile (kernels_dense.cu)
  double* gemv_partials = nullptr; // [ksplit][tiles][rows per tile] f64 (owned by the plan)
  int acc = ACC_SET;
  int64_t fill_len = 0;
  // ST_GATHER: dst[i] (acc)= src[idx[i]] (g_scatter=0) or dst[idx[i]] (acc)= src[i] (g_scatter=1), i < g_n
  const void* g_idx = nullptr;
  int64_t g_n = 0;
  int g_scatter = 0, g_idx64 = 0;
  // ST_GEMM_TC (kernels_gemm_tc.cu): pre-split right-hand sides, tensor maps, tile tables
  int32_t tc_np = 0, tc_nsegs = 0;
  int64_t tc_kp = 0;
  void* tc_xs = nullptr;
  void* tc_maps = nullptr;
  void* tc_segs = nullptr;
  int32_t* tc_tile_ptr = nullptr;
  int32_t* tc_koff = nullptr;
  alignas(64) unsigned char tc_xmap[128] = {};
};

struct Plan {
  std::vector<Step> steps;
  std::vector<void*> tmps;       // device temporaries (owned)
  std::vector<size_t> tmp_bytes;
  std::vector<void*> blobs;      // device table blobs (owned)
  int engines = 0;
  // Validity: a plan holds raw pointers to the linearization points (mo) of the pointwise leaves it evaluated.
  // It stays valid while every one of those leaves still points at the same buffer (`points`); trees with more
  // than kMaxTrackedPoints such leaves fall back to the global point! epoch (`version`).  Plans that read no
  // linearization point (every linear operator) are never invalidated by anybody's point!.
  static constexpr size_t kMaxTrackedPoints = 64;
  std::vector<std::pair<jets_op, const void*>> points;
  std::vector<jets_buf> point_bufs;   // those buffers, retained: a plan pinned by a captured CUDA graph outlives a later point!
  bool uses_point = false;
  uint64_t version = 0;
  bool valid() const;
  ~Plan();
};

// plan.cu
// Distributed block-banded apply (dist.cu): which rows of the rank-local operator one launch covers.
// Flag words a launch may wait on (in THIS rank's arena) ...
enum : int { GF_LO_READY = 0, GF_HI_READY = 1, GF_PREV_DONE = 2, GF_NEXT_DONE = 3 };
// ... and the signals it may raise (flag words in the NEIGHBOURS' arenas): previous rank's HI_READY, next
// rank's LO_READY, previous rank's NEXT_DONE, next rank's PREV_DONE.
enum : int { GS_PREV_HI_READY = 0, GS_NEXT_LO_READY = 1, GS_PREV_NEXT_DONE = 2, GS_NEXT_PREV_DONE = 3 };
struct BandedSel {
  int mode = JETS_MODE_DF;
  int row_begin = 0, row_end = 0;               // forward: local block rows; adjoint: own block columns
  bool send_prev = false, send_next = false;    // include the push rows (forward) / partial-sum rows (adjoint)
  bool has_prev = false, has_next = false;
  int owned = 0;                                // signals this launch must raise even if no unit feeds them
  bool pull = false;                            // forward only: halo terms read the neighbours' vectors in place
};
std::shared_ptr<Plan> build_banded_plan(jets_op A_loc, int halo, const BandedSel& sel);
std::shared_ptr<Plan> get_plan(jets_op a, int mode, int accumulate);
void run_plan(Plan& p, int dtype, char* in, char* out);
struct ApplyCoef;
bool run_plan_axpby(Plan& p, int dtype, char* in, char* out, const ApplyCoef& coef);  // false: plan is not one bundle launch

// kernels_fused.cu
void launch_fused(const DevFused& f, int dtype, const char* in, char* out, cudaStream_t s);
int fused_tile_elems(int dtype);
int fused_nslots(int slot_streams);
int fast_tile_bytes(int variant);
int fast_nslots(int variant, int slot_streams);
void launch_fused_fast(const DevFused& f, int dtype, const char* in, char* out, cudaStream_t s);
// kernels_fused_bundle.cu
int bundle_buf_bytes(int variant);
int bundle_smem_budget();
// out = cA * (A in) + cO * out_old in the store epilogue of the bundle kernel; each coefficient is a
// device scalar (with JETS_COEF_* flags applied) or a constant.
struct ApplyCoef {
  const double* a_ptr = nullptr; double a_const = 1.0; int a_flags = 0;
  const double* o_ptr = nullptr; double o_const = 0.0; int o_flags = 0;
};
// Per-launch cross-rank wiring of a gated bundle launch: flag words in THIS rank's exchange arena that
// neighbours raise (a unit whose bundle names flag k starts only once flags[k*stride] >= wait_val[k]),
// flag words in the NEIGHBOURS' arenas this launch raises to sig_val[k] once every unit feeding signal k
// has been stored, and the alternative input / output bases (halo and staging buffers, local or peer).
struct GateLaunch {
  const uint32_t* flags = nullptr;
  uint32_t wait_val[kGateFlags] = {0, 0, 0, 0};
  uint32_t* sig_addr[kGateFlags] = {nullptr, nullptr, nullptr, nullptr};
  uint32_t sig_val[kGateFlags] = {0, 0, 0, 0};
  const char* in_alt[3] = {nullptr, nullptr, nullptr};
  char* out_alt[3] = {nullptr, nullptr, nullptr};
  // escape hatch: a wait that lasts longer than timeout_ns (0 = forever) bumps *err and proceeds with whatever is
  // in the buffer -- a missing neighbour must never hang the GPU; the host turns err != 0 into an error
  uint32_t* err = nullptr;
  uint64_t timeout_ns = 0;
  // flag words that must have reached exit_val before the LAST CTA of the launch may exit (pull mode: the
  // neighbours have finished reading this rank's input vector, so whatever follows on the stream may overwrite it)
  int32_t exit_wait = 0;
  uint32_t exit_val[kGateFlags] = {0, 0, 0, 0};
};
void launch_fused_bundle(const DevFused& f, int dtype, const char* in, char* out, cudaStream_t s,
                         const ApplyCoef* coef = nullptr, const GateLaunch* gate = nullptr);

// kernels_dense.cu
void launch_gemv(const Step& st, int dtype, const char* in, char* out, cudaStream_t s);
void gemv_tile_count(int dtype, bool trans, int32_t out_len, int32_t* ntiles);

// kernels_gemm_tc.cu
bool gemm_tc_eligible(int dtype, const std::vector<DBlock>& blocks);
void gemm_tc_prepare(Step& st, Plan& plan);
void launch_gemm_tc(const Step& st, const char* in, char* out, cudaStream_t s);

// kernels_vec.cu
void vec_fill(int dtype, void* p, int64_t n, double a, cudaStream_t s);
void vec_rand(int dtype, void* p, int64_t n, uint64_t seed, uint64_t off, int dist, cudaStream_t s);
void vec_lincomb(int dtype, void* out, int64_t n, int k, const double* c, const void* const* x,
                 cudaStream_t s);
void vec_hadamard(int dtype, void* out, const void* x, const void* y, int64_t n, cudaStream_t s);
// reductions: kind 0 dot, 1 sumsq, 2 sumabs, 3 count nonzero, 4 max|x|, 5 min|x|, 6 sum|x|^p,
// 7 min, 8 max.  Result (f64) is written to dev_out (device pointer).
void vec_reduce(int dtype, int kind, const void* x, const void* y, int64_t n, double p,
                double* dev_out, cudaStream_t s);
void scalar_op(double* out, char op, const double* a, const double* b, cudaStream_t s);
constexpr int kMaxScalarProg = 16;
struct ScalarProg {
  double* out[kMaxScalarProg]; const double* a[kMaxScalarProg]; const double* b[kMaxScalarProg]; char op[kMaxScalarProg]; int n;
  // filled in by scalar_prog(): which earlier step produces an operand (-1: read it from memory) and whether a step is
  // the last one that writes its scalar -- so the kernel loads every memory operand at once, evaluates the steps from
  // registers / shared memory and stores once (11 dependent global round trips took 6 us inside the LSQR iteration)
  int8_t asrc[kMaxScalarProg], bsrc[kMaxScalarProg]; uint8_t store[kMaxScalarProg];
};
void scalar_prog(const ScalarProg& p, cudaStream_t s);
void vec_axpby_dev(int dtype, void* out, int64_t n, const double* sa, double ca, int af,
                   const void* x, const double* sb, double cb, int bf, const void* y,
                   cudaStream_t s);
void vec_axpby_pair_dev(int dtype, int64_t n, void* const out[2], const double* const sa[2], const double ca[2], const int af[2],
                        const void* const x[2], const double* const sb[2], const double cb[2], const int bf[2], const void* const y[2],
                        cudaStream_t s);
void scalar_finish_norm(double* v, double p, cudaStream_t s);
// restriction / its adjoint: out[i] (acc)= in[idx[i]]  or  out[idx[i]] (acc)= in[i]   (indices unique)
void vec_gather(int dtype, void* out, const void* in, const void* idx, int idx64, int64_t n, int scatter, int acc, cudaStream_t s);

// kernels_cplx.cu: the same vector-space kernels for the complex eltypes (n counts complex elements)
void cvec_fill(int dtype, void* p, int64_t n, double re, double im, cudaStream_t s);
void cvec_lincomb(int dtype, void* out, int64_t n, int k, const double* c_re_im, const void* const* x,
                  cudaStream_t s);
void cvec_hadamard(int dtype, void* out, const void* x, const void* y, int64_t n, int conj_x, cudaStream_t s);
void cvec_abs(int dtype, void* out_real, const void* x, int64_t n, cudaStream_t s);
// kind: 0 Re<x,y>, 1 Im<x,y> (<x,y> = sum conj(x) y), 2 sum w|x|^2, 3 sum w|x|, 4 sum w[x!=0],
// 5 max|x|, 6 min|x|, 7 sum w|x|^p.  w (f64 weights, one per element) may be null; finish as vec_reduce.
void cvec_reduce(int dtype, int kind, const void* x, const void* y, const double* w, int64_t n, double p,
                 int finish, double* dev_out, cudaStream_t s);

inline void count_launch(int n = 1) { ctx().launches += n; }

// Programmatic dependent launch for the small vector kernels (reductions, axpby, scalar programs: the kernels between
// the applies of a solver iteration) -- OFF by default (JETS_B200_VEC_PDL=1): inside the captured LSQR iteration it
// measured 1.5-5 % SLOWER than plain kernel nodes (profiles/r02_ab_vec_pdl.log), without PDL griddepcontrol.* are
// no-ops.  Every such kernel starts with pdl_enter(): griddepcontrol.wait FIRST, then
// griddepcontrol.launch_dependents -- so the next kernel of the stream is launched (and runs its own prologue up to
// its wait) while this one does its work, but never while the kernel BEFORE this one is still running: at most two
// consecutive kernels of the stream overlap, which is what the "operator state before the wait" rule of the bundle
// kernel (Context::pdl_out_lo/hi = what the previous launch writes) assumes.  `wlo`/`wbytes`: what the launch writes.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, cudaStream_t s, const void* wlo, size_t wbytes,
                       Args&&... args) {
  Context& c = ctx();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (c.no_pdl || !c.vec_pdl) ? 0 : 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
  c.pdl_out_lo = reinterpret_cast<uintptr_t>(wlo);
  c.pdl_out_hi = c.pdl_out_lo + wbytes;
}
#endif

}  // namespace jets
