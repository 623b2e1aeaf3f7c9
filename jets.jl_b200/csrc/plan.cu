// Lowers an operator tree (Jet) + mode into a short list of kernel launches, cached per handle.
//
//   * Every subtree made of elementwise / stencil leaves under JopBlock, JopSum, JopComposite,
//     JopAdjoint is flattened into ONE fused launch: per output block a list of signed terms,
//     each an elementwise chain over one input block (src/Jets.jl:524-540, 630-655, 1010-1057
//     collapse into a table walk; no temporaries, no zero-fill + accumulate passes).
//   * Dense leaves (fusion barriers) become block-GEMV launches over a table of matrices.
//   * Anything else is staged through device temporaries, stage by stage, exactly the way the
//     reference evaluates it (zeros(range(op)) per composite stage, :525-539).
#include <algorithm>
#include <cstdlib>
#include "common.hpp"
#include "cplx.cuh"

namespace jets {

uint64_t g_epoch = 1;  // bumped by point! -> cached plans holding stale mo pointers are rebuilt

namespace {

constexpr int kMaxStagesPerTerm = 6;
constexpr int kMaxStreamsPerTerm = 4;
constexpr size_t kMaxEntries = 1u << 20;
constexpr int kMaxRing = 16;   // x ring / state ring capacity of the bundle kernel

struct Entry {
  int r = 0, c = 0;
  int sign = 1;
  bool linear = true;
  std::vector<FStage> chain;  // application order
};
using Entries = std::vector<Entry>;

int map_mode_lnview(int mode) { return mode == JETS_MODE_F ? JETS_MODE_DF : mode; }
int map_mode_adj(int mode) { return mode == JETS_MODE_DFT ? JETS_MODE_DF : JETS_MODE_DFT; }

const Space& out_space(jets_op a, int mode) { return mode == JETS_MODE_DFT ? a->dom : a->rng; }
const Space& in_space(jets_op a, int mode) { return mode == JETS_MODE_DFT ? a->rng : a->dom; }

// c1 != 0 marks an operand stream that lives in caller-owned memory WITHOUT the 256-byte guard of library
// allocations (jets_buf_wrap): the TMA engines fetch 16 bytes before a tile (stencil halo) and round its
// tail up, so such streams are only ever read by the guarded-load (LDG) engine.
FStage mk(int op, int fn = 0, const void* ptr = nullptr, double c0 = 0, bool guarded = true) {
  FStage s;
  s.op = op; s.fn = fn; s.ptr = ptr; s.c0 = c0; s.c1 = guarded ? 0.0 : 1.0;
  return s;
}
bool stream_tma_ok(const FStage& s) { return !s.ptr || ((reinterpret_cast<uintptr_t>(s.ptr) & 15) == 0 && s.c1 == 0.0); }

int chain_streams(const std::vector<FStage>& ch) {
  int n = 1;
  for (auto& s : ch) n += s.ptr != nullptr;
  return n;
}

bool expand(jets_op a, int mode, Entries& out, Space& isp, Space& osp);
thread_local size_t g_esz = 8;  // element size of the plan being built
thread_local std::vector<std::pair<jets_op, const void*>>* g_points = nullptr;  // linearization points the plan being built reads
thread_local bool g_cplx = false;  // complex eltype: adjoint stages conjugate their operand stream

bool has_stencil(const Entry& e) {
  for (auto& s : e.chain)
    if (s.op == S_FDIFF || s.op == S_BDIFF || s.op == S_LAP) return true;
  return false;
}

// Entries of an operator acting on ONE flat block are re-expressed on the block structure
// `target` (same total length): an elementwise chain acts block by block, its operand streams
// shifted by the block offset.  (The reference applies e.g. `a*A`'s scalar stage to a BlockArray
// through the per-block broadcast, src/Jets.jl:905-911, :1159-1164.)
bool reblock(Entries& es, const Space& target) {
  for (auto& e : es)
    if (e.r != 0 || e.c != 0 || has_stencil(e)) return false;
  Entries out;
  int64_t off = 0;
  for (size_t k = 0; k < target.len.size(); ++k) {
    for (const Entry& e : es) {
      Entry x = e;
      x.r = x.c = (int)k;
      for (auto& st : x.chain)
        if (st.ptr) st.ptr = reinterpret_cast<const char*>(st.ptr) + (size_t)off * g_esz;
      out.push_back(std::move(x));
    }
    off += target.len[k];
  }
  es.swap(out);
  return true;
}

bool is_flat(const Space& s) { return s.len.size() == 1; }

// B is applied first, then A:  result = A after B.
bool product(const Entries& A, const Entries& B, Entries& out) {
  for (const Entry& a : A) {
    int nb = 0;
    for (const Entry& b : B) nb += (b.r == a.c);
    if (nb > 1 && !a.linear) return false;  // phi(sum) != sum(phi)
    if (nb == 0) {
      if (a.linear) continue;  // A(0) = 0
      Entry e;
      e.r = a.r; e.c = 0; e.sign = a.sign; e.linear = false;
      e.chain.push_back(mk(S_ZERO));
      e.chain.insert(e.chain.end(), a.chain.begin(), a.chain.end());
      out.push_back(std::move(e));
      continue;
    }
    for (const Entry& b : B) {
      if (b.r != a.c) continue;
      Entry e;
      e.r = a.r; e.c = b.c;
      e.linear = a.linear && b.linear;
      e.chain = b.chain;
      if (b.sign < 0 && !a.linear) {
        e.chain.push_back(mk(S_NEG));
        e.sign = a.sign;
      } else {
        e.sign = a.sign * b.sign;
      }
      e.chain.insert(e.chain.end(), a.chain.begin(), a.chain.end());
      if ((int)e.chain.size() > kMaxStagesPerTerm) return false;
      if (chain_streams(e.chain) > kMaxStreamsPerTerm) return false;
      out.push_back(std::move(e));
      if (out.size() > kMaxEntries) return false;
    }
  }
  return true;
}

// ops given in APPLICATION order (first applied first), all expanded with `mode`.
bool expand_seq(const std::vector<jets_op>& seq, int mode, Entries& out, Space& isp, Space& osp) {
  Entries cur;
  Space cur_in, cur_out;
  if (!expand(seq[0], mode, cur, cur_in, cur_out)) return false;
  for (size_t i = 1; i < seq.size(); ++i) {
    Entries nxt, prod;
    Space nin, nout;
    if (!expand(seq[i], mode, nxt, nin, nout)) return false;
    if (!cur_out.same_layout(nin)) {
      if (cur_out.total() != nin.total()) return false;
      if (is_flat(nin) && is_flat(nout) && nin.total() == nout.total()) {
        if (!reblock(nxt, cur_out)) return false;
        nin = nout = cur_out;
      } else if (is_flat(cur_out) && is_flat(cur_in) && cur_in.total() == cur_out.total()) {
        if (!reblock(cur, nin)) return false;
        cur_in = cur_out = nin;
      } else {
        return false;
      }
    }
    if (!product(nxt, cur, prod)) return false;
    cur.swap(prod);
    cur_out = nout;
  }
  out.swap(cur);
  isp = cur_in;
  osp = cur_out;
  return true;
}

std::vector<jets_op> app_order(jets_op a, int mode) {
  std::vector<jets_op> seq;
  if (mode == JETS_MODE_DFT) seq.assign(a->kids.begin(), a->kids.end());       // ops[1]' first (:537)
  else seq.assign(a->kids.rbegin(), a->kids.rend());                          // ops[n] first (:525)
  return seq;
}

bool expand(jets_op a, int mode, Entries& out, Space& isp, Space& osp) {
  isp = in_space(a, mode);
  osp = out_space(a, mode);
  switch (a->kind) {
    case K_DIAG: {
      Entry e;
      // df'!: m .= conj(w) .* d (fixture JopFoo, test/runtests.jl:4); conj is the identity on reals
      e.chain.push_back(mk(S_DIAG, (g_cplx && mode == JETS_MODE_DFT) ? kConjFlag : 0, a->w->ptr(), 0, a->w->guarded()));
      out.push_back(e);
      return true;
    }
    case K_SCALE: {
      Entry e;
      if (g_cplx && a->ai != 0.0) {   // _constdiag_df'!: m .= conj(a) .* d (src/Jets.jl:1160)
        e.chain.push_back(mk(S_CSCALE, 0, nullptr, a->a));
        e.chain.push_back(mk(S_CIMAG, 0, nullptr, mode == JETS_MODE_DFT ? -a->ai : a->ai));
      } else {
        e.chain.push_back(mk(S_SCALE, 0, nullptr, a->a));
      }
      out.push_back(e);
      return true;
    }
    case K_PW: {
      Entry e;
      if (mode == JETS_MODE_F) {
        e.linear = false;
        e.chain.push_back(mk(S_PW_F, a->fn, nullptr, a->p));
      } else {
        JETS_CHECK(a->mo != nullptr, JETS_ERR_NO_POINT,
                   "Jacobian of a pointwise operator applied before point!/jacobian set mo");
        if (g_points) g_points->push_back({a, a->mo->ptr()});
        e.chain.push_back(mk(S_PW_J, a->fn | ((g_cplx && mode == JETS_MODE_DFT) ? kConjFlag : 0), a->mo->ptr(), a->p, a->mo->guarded()));
      }
      out.push_back(e);
      return true;
    }
    case K_STENCIL: {
      Entry e;
      if (a->fn == JETS_ST_FDIFF) e.chain.push_back(mk(mode == JETS_MODE_DFT ? S_BDIFF : S_FDIFF));
      else e.chain.push_back(mk(S_LAP));
      out.push_back(e);
      return true;
    }
    case K_ZERO: return true;  // contributes nothing (src/Jets.jl:942, skipped at :1022,:1047)
    case K_DENSE: return false;
    case K_RESTRICT: return false;   // changes the length: a fusion barrier with its own gather/scatter kernel
    case K_LNVIEW: return expand(a->kids[0], map_mode_lnview(mode), out, isp, osp);
    case K_ADJ: return expand(a->kids[0], map_mode_adj(mode), out, isp, osp);
    case K_COMPOSE: return expand_seq(app_order(a, mode), mode, out, isp, osp);
    case K_SUM: {
      bool first = true;
      for (size_t k = 0; k < a->kids.size(); ++k) {
        Entries e;
        Space ki, ko;
        if (!expand(a->kids[k], mode, e, ki, ko)) return false;
        if (first) { isp = ki; osp = ko; first = false; }
        else if (!ki.same_layout(isp) || !ko.same_layout(osp)) {
          if (is_flat(ki) && is_flat(ko) && isp.same_layout(osp)) {
            if (!reblock(e, isp)) return false;
          } else if (is_flat(isp) && is_flat(osp) && ki.same_layout(ko)) {
            if (!reblock(out, ki)) return false;
            isp = ki; osp = ko;
          } else {
            return false;
          }
        }
        for (auto& x : e) {
          x.sign *= a->sgn[k];
          out.push_back(std::move(x));
        }
      }
      return true;
    }
    case K_BLOCK: {
      for (int c = 0; c < a->C; ++c)
        for (int r = 0; r < a->R; ++r) {
          Entries e;
          Space ki, ko;
          if (!expand(a->kids[r + (size_t)c * a->R], mode, e, ki, ko)) return false;
          if (!is_flat(ki) || !is_flat(ko)) return false;  // nested block spaces are not supported
          for (auto& x : e) {
            if (x.r != 0 || x.c != 0) return false;
            if (mode == JETS_MODE_DFT) { x.r = c; x.c = r; }
            else { x.r = r; x.c = c; }
            out.push_back(std::move(x));
            if (out.size() > kMaxEntries) return false;
          }
        }
      return true;
    }
  }
  return false;
}

void halo_of(const Entries& es, int& hl, int& hr) {
  hl = hr = 0;
  for (const Entry& e : es) {
    int l = 0, r = 0;
    for (auto it = e.chain.rbegin(); it != e.chain.rend(); ++it) {
      if (it->op == S_FDIFF) r += 1;
      else if (it->op == S_BDIFF) l += 1;
      else if (it->op == S_LAP) { l += 1; r += 1; }
    }
    hl = std::max(hl, l);
    hr = std::max(hr, r);
  }
}

// Which straight-line fast path (if any) evaluates this chain.
int classify(const std::vector<FStage>& ch) {
  auto is = [&](size_t i, int op) { return i < ch.size() && ch[i].op == op; };
  auto j2 = [&](size_t i) { return is(i, S_PW_J) && ch[i].fn == JETS_PW_SQUARE; };
  const size_t n = ch.size();
  if (n == 0) return PAT_COPY;
  if (n == 1) {
    if (is(0, S_DIAG)) return PAT_DIAG;
    if (is(0, S_FDIFF)) return PAT_FDIFF;
    if (is(0, S_BDIFF)) return PAT_BDIFF;
    if (is(0, S_LAP)) return PAT_LAP;
    if (is(0, S_SCALE)) return PAT_SCALE;
    if (j2(0)) return PAT_J2;
    if (is(0, S_PW_F) && ch[0].fn == JETS_PW_SQUARE) return PAT_SQUARE;
  }
  if (n == 2 && is(1, S_SCALE)) {
    if (is(0, S_LAP)) return PAT_LAP_SCALE;
    if (is(0, S_FDIFF)) return PAT_FDIFF_SCALE;
    if (is(0, S_BDIFF)) return PAT_BDIFF_SCALE;
  }
  if (n == 2 && is(0, S_SCALE)) {
    if (is(1, S_LAP)) return PAT_SCALE_LAP;
    if (is(1, S_FDIFF)) return PAT_SCALE_FDIFF;
    if (is(1, S_BDIFF)) return PAT_SCALE_BDIFF;
  }
  auto plain_diag = [&](size_t i) { return is(i, S_DIAG) && ch[i].fn == 0; };   // real diagonal (no conjugation flag)
  if (n == 2 && plain_diag(1)) {
    if (is(0, S_LAP)) return PAT_LAP_DIAG;
    if (is(0, S_FDIFF)) return PAT_FDIFF_DIAG;
    if (is(0, S_BDIFF)) return PAT_BDIFF_DIAG;
  }
  if (n == 2 && plain_diag(0)) {
    if (is(1, S_LAP)) return PAT_DIAG_LAP;
    if (is(1, S_FDIFF)) return PAT_DIAG_FDIFF;
    if (is(1, S_BDIFF)) return PAT_DIAG_BDIFF;
  }
  if (n == 3 && j2(0) && is(1, S_FDIFF) && is(2, S_DIAG)) return PAT_J2_FDIFF_DIAG;
  if (n == 3 && is(0, S_DIAG) && is(1, S_BDIFF) && j2(2)) return PAT_DIAG_BDIFF_J2;
  return PAT_GENERIC;
}

std::vector<int64_t> offsets_of(const Space& s) {
  std::vector<int64_t> o(s.len.size() + 1, 0);
  for (size_t i = 0; i < s.len.size(); ++i) o[i + 1] = o[i] + s.len[i];
  return o;
}

// ------------------------------------------------------------------ plan builder ---------
struct Builder {
  Plan& plan;
  int dtype;
  bool io_ok;     // apply's in/out are library-owned (guarded) and 16B aligned
  int engine;     // 0 auto, 1 TMA, 2 LDG

  Ref new_tmp(int64_t elems) {
    const size_t bytes = (size_t)elems * dsize(dtype);
    char* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, bytes + 2 * kGuardBytes));
    // on the context stream (non-blocking: the legacy stream's memset would not be ordered against it)
    CUDA_TRY(cudaMemsetAsync(p, 0, bytes + 2 * kGuardBytes, ctx().stream));
    plan.tmps.push_back(p);
    plan.tmp_bytes.push_back(bytes);
    Ref r;
    r.which = 2 + (int)plan.tmps.size() - 1;
    r.off = 0;
    return r;
  }
  bool ref_ok(const Ref& r) const { return r.which >= 2 || io_ok; }

  void emit_fill0(Ref dst, int64_t len) {
    Step st;
    st.kind = ST_FILL0;
    st.dst = dst;
    st.fill_len = len;
    plan.steps.push_back(std::move(st));
  }

  // Entries use block coordinates of (out_sp, in_sp); dst/src give the flat base offsets.
  void emit_fused(Entries& es, const Space& out_sp, const Space& in_sp, Ref dst, Ref src, int acc) {
    int hl, hr;
    halo_of(es, hl, hr);
    JETS_CHECK(hl <= 1 && hr <= 1, JETS_ERR_UNSUPPORTED, "internal: halo too wide for fusion");
    std::stable_sort(es.begin(), es.end(), [](const Entry& x, const Entry& y) {
      return x.r != y.r ? x.r < y.r : x.c < y.c;
    });
    const auto oo = offsets_of(out_sp), io = offsets_of(in_sp);
    const size_t esz = dsize(dtype);
    FusedTables t;
    t.hl = hl; t.hr = hr;
    // Pre-pass: can the TMA engines be used at all (every stream 16B aligned in guarded memory),
    // does every chain have a straight-line fast path, and how many terms does a row sum?
    bool tma = ref_ok(dst) && ref_ok(src) && (dst.off * esz) % 16 == 0 && (src.off * esz) % 16 == 0;
    bool all_fast = !ctx().no_fast && engine != 2;
    int max_row_terms = 0, max_term_streams = 1;
    {
      int run = 0, prev_r = -1;
      for (const Entry& e : es) {
        if (classify(e.chain) == PAT_GENERIC) all_fast = false;
        for (const FStage& s : e.chain)
          if (!stream_tma_ok(s)) tma = false;
        if (((src.off + io[e.c]) * esz) % 16 || ((dst.off + oo[e.r]) * esz) % 16) tma = false;
        run = (e.r == prev_r) ? run + 1 : 1;
        prev_r = e.r;
        max_row_terms = std::max(max_row_terms, run);
        max_term_streams = std::max(max_term_streams, chain_streams(e.chain));
      }
      for (size_t r = 0; r < out_sp.len.size(); ++r)
        if (((dst.off + oo[r]) * esz) % 16) tma = false;
    }
    if (engine == 2) tma = false;
    all_fast = all_fast && tma && !is_cplx(dtype);    // complex spaces: the interpreter kernels (TMA-staged or guarded loads)
    // Fast-kernel shape (measured on B200, profiles/README.md): 16 consumer warps x 2 vectors per
    // thread (16 KB tiles) amortises the per-slot work best -- 81% of HBM peak on config 5, 98% on
    // config 2 vs 53% / 87% with 8 KB tiles.  Blocks shorter than a few tiles keep 8 KB tiles.
    int variant = ctx().fast_variant;
    (void)max_row_terms; (void)max_term_streams;
    if (variant < 0 || variant > 5) {
      int64_t longest = 0;
      for (auto l : out_sp.len) longest = std::max(longest, l);
      variant = (longest * (int64_t)esz >= 4 * 16384) ? 2 : 0;
    }
    if (all_fast && !ctx().no_bundle && engine != 3 && emit_bundle(es, out_sp, in_sp, dst, src, acc, hl, hr, variant)) return;
    const int tile = all_fast ? fast_tile_bytes(variant) / (int)esz : fused_tile_elems(dtype);
    size_t k = 0;
    bool heavy = false;
    constexpr int kSlotStreams = 4, kGroupTermsMax = 4, kGroupStagesMax = 12;
    for (size_t r = 0; r < out_sp.len.size(); ++r) {
      FRow row{};
      row.out_off = dst.off + oo[r];
      row.len = out_sp.len[r];
      row.term_begin = (int32_t)t.terms.size();
      row.group_begin = (int32_t)t.groups.size();
      GroupRec grp{};
      int g_terms = 0, g_stages = 0;
      auto close_group = [&]() {
        if (g_terms == 0) return;
        grp.nterms = g_terms;
        t.max_streams = std::max(t.max_streams, (int)grp.nstreams);
        t.groups.push_back(grp);
        grp = GroupRec{};
        g_terms = g_stages = 0;
      };
      while (k < es.size() && es[k].r == (int)r) {
        const Entry& e = es[k++];
        FTerm tm{};
        tm.in_off = src.off + io[e.c];
        tm.in_abs = nullptr;
        tm.stage_begin = (int32_t)t.stages.size();
        for (const FStage& s : e.chain) {
          t.stages.push_back(s);
          if (!stream_tma_ok(s)) tma = false;
          if ((s.op == S_PW_F || s.op == S_PW_J) && (s.fn & ~kConjFlag) != JETS_PW_SQUARE) heavy = true;
        }
        tm.stage_end = (int32_t)t.stages.size();
        tm.sign = (acc == ACC_SUB) ? -e.sign : e.sign;
        tm.nstreams = chain_streams(e.chain);
        if ((tm.in_off * esz) % 16) tma = false;
        JETS_CHECK(in_sp.len[e.c] == out_sp.len[r], JETS_ERR_SHAPE,
                   "elementwise block (%d,%d) maps %lld -> %lld elements", (int)r, e.c,
                   (long long)in_sp.len[e.c], (long long)out_sp.len[r]);
        t.terms.push_back(tm);
        // close the current group when this term does not fit into its slot any more
        if (g_terms > 0 && (grp.nstreams + tm.nstreams > kSlotStreams || g_terms + 1 > kGroupTermsMax ||
                            g_stages + (int)e.chain.size() > kGroupStagesMax))
          close_group();
        GTerm gt{};
        gt.stage0 = (uint8_t)g_stages;
        gt.nstages = (uint8_t)e.chain.size();
        gt.stream0 = (uint8_t)grp.nstreams;
        gt.sign = (int8_t)tm.sign;
        gt.pattern = (uint8_t)classify(e.chain);
        grp.terms[g_terms++] = gt;
        grp.ptr[grp.nstreams] = (int64_t)(tm.in_off * (int64_t)esz);   // relative to the apply's `in`
        grp.rel_mask |= 1 << grp.nstreams;
        grp.nstreams++;
        for (const FStage& s : e.chain) {
          CStage cs{};
          cs.op = (uint8_t)s.op; cs.fn = (uint8_t)s.fn; cs.has_stream = s.ptr != nullptr; cs.c0 = s.c0;
          grp.stages[g_stages++] = cs;
          if (s.ptr) grp.ptr[grp.nstreams++] = (int64_t)reinterpret_cast<uintptr_t>(s.ptr);
        }
      }
      row.term_end = (int32_t)t.terms.size();
      close_group();
      row.group_end = (int32_t)t.groups.size();
      row.init = (acc == ACC_SET) ? 0 : 1;
      row.ntiles = (int32_t)((row.len + tile - 1) / tile);
      if ((row.out_off * esz) % 16) tma = false;
      if (row.term_end == row.term_begin && acc != ACC_SET) continue;  // nothing to add
      if (row.len == 0) continue;
      t.rows.push_back(row);
    }
    if ((all_fast ? fast_nslots(variant, t.max_streams) : fused_nslots(t.max_streams)) < 2) tma = false;
    t.tma_ok = tma;
    bool use_tma = tma;
    if (engine == 2) use_tma = false;
    if (engine == 1 && !tma) JETS_FAIL(JETS_ERR_UNSUPPORTED, "TMA engine forced but operands are not aligned/guarded");

    // schedule: tiles are visited in (super-chunk of S tiles, row, tile) order and dealt round-robin
    // to the CTAs.  Rows sorted by super-chunk count (desc); one segment per constant set of active rows.
    int64_t max_tiles = 1;
    for (auto& r : t.rows) max_tiles = std::max<int64_t>(max_tiles, r.ntiles);
    const int K = (int)std::min<int64_t>(max_tiles, 4 * (int64_t)ctx().sm_count);
    std::vector<int32_t> order(t.rows.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
    auto nchunks = [&](int32_t i) { return (int64_t)(t.rows[i].ntiles + K - 1) / K; };
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return nchunks(x) > nchunks(y); });
    std::vector<FSeg> segs;
    int64_t pos = 0, begin = 0, ntiles = 0;
    for (auto& r : t.rows) ntiles += r.ntiles;
    for (int kact = (int)order.size(); kact >= 1; --kact) {
      const int64_t upper = nchunks(order[kact - 1]);
      if (upper > pos) {
        FSeg sg{};
        sg.tile_begin = begin; sg.pos_begin = pos; sg.nactive = kact;
        segs.push_back(sg);
        begin += (upper - pos) * kact;
        pos = upper;
      }
    }
    Step st;
    st.kind = ST_FUSED;
    st.src = src; st.dst = dst; st.acc = acc;
    DevFused& f = st.fused;
    f.nrows = (int32_t)t.rows.size();
    f.nsegs = (int32_t)segs.size();
    f.ntiles = ntiles;
    f.nitems = begin;
    f.hl = hl; f.hr = hr; f.slot_streams = t.max_streams;
    f.S = K;
    f.heavy = heavy;
    f.fast = all_fast && use_tma;
    f.variant = variant;
    f.tile_elems = tile;
    f.use_tma = use_tma;
    if (f.nrows > 0) {
      auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
      const size_t b0 = 0, b1 = b0 + al(t.stages.size() * sizeof(FStage)),
                   b2 = b1 + al(t.terms.size() * sizeof(FTerm)), b3 = b2 + al(t.rows.size() * sizeof(FRow)),
                   b4 = b3 + al(segs.size() * sizeof(FSeg)), b5x = b4 + al(order.size() * sizeof(int32_t)),
                   b5 = b5x + al(t.groups.size() * sizeof(GroupRec));
      std::vector<char> host(b5, 0);
      if (!t.stages.empty()) memcpy(host.data() + b0, t.stages.data(), t.stages.size() * sizeof(FStage));
      if (!t.terms.empty()) memcpy(host.data() + b1, t.terms.data(), t.terms.size() * sizeof(FTerm));
      memcpy(host.data() + b2, t.rows.data(), t.rows.size() * sizeof(FRow));
      memcpy(host.data() + b3, segs.data(), segs.size() * sizeof(FSeg));
      memcpy(host.data() + b4, order.data(), order.size() * sizeof(int32_t));
      if (!t.groups.empty()) memcpy(host.data() + b5x, t.groups.data(), t.groups.size() * sizeof(GroupRec));
      char* blob = nullptr;
      CUDA_TRY(cudaMalloc(&blob, b5));
      CUDA_TRY(cudaMemcpy(blob, host.data(), b5, cudaMemcpyHostToDevice));
      plan.blobs.push_back(blob);
      f.blob = blob;
      f.stages = reinterpret_cast<FStage*>(blob + b0);
      f.terms = reinterpret_cast<FTerm*>(blob + b1);
      f.rows = reinterpret_cast<FRow*>(blob + b2);
      f.segs = reinterpret_cast<FSeg*>(blob + b3);
      f.order = reinterpret_cast<int32_t*>(blob + b4);
      f.groups = reinterpret_cast<GroupRec*>(blob + b5x);
    }
    plan.engines |= use_tma ? 1 : 2;
    plan.steps.push_back(std::move(st));
  }

  // ---- bundle engine (kernels_fused_bundle.cu) ------------------------------------------------
  // Rows are grouped into BUNDLES of consecutive equal-length rows that share input blocks; one CTA
  // evaluates a whole bundle at one tile position, keeping every input tile in the shared-memory x
  // ring from its first to its last use.  The planner replays the kernel's ring allocation (ring
  // order, `NX` buffers) to fix, per term, the allocation index of its input tile and whether the
  // term is the tile's first use (the producer loads it) and/or its last use (consumers release it).
  struct PTerm {
    int pattern = 0, sign = 1;
    int64_t key = 0;                 // input block: byte offset from its base pointer
    int in_alt = 0;                  // base: 0 = the apply's `in`, k = GateLaunch::in_alt[k-1]
    std::vector<int64_t> sptr;       // state streams
    std::vector<CStage> stages;
  };
  struct PRow {
    int64_t out_off = 0, len = 0;    // out_off: element offset from the row's output base
    int out_alt = 0;                 // base: 0 = the apply's `out`, k = GateLaunch::out_alt[k-1]
    int wait = 0, sig = 0;           // cross-rank gate: flag words to wait for / signals the finished row feeds
    bool early = false;              // nothing but a store to peer memory: claimed by a few CTAs from the early queue
    std::vector<PTerm> terms;
  };
  struct BundleSim {
    std::vector<BGroupRec> groups;
    std::vector<BundleRec> bundles;
    std::vector<char> early;         // per bundle: every row is an early row
    std::vector<std::pair<int, int>> row_range;   // per bundle: [first, last) of the launch's row list
    int maxdist = 0, sstreams = 0, max_rows = 1;
  };

  // state streams one term group (= one state slot) may carry; JETS_B200_GROUP_STREAMS (tuning): fewer streams per
  // slot leave room for 16 KB tiles when every term has a state stream (config 1)
  static int max_group_streams() {
    const int v = ctx().group_streams;
    return (v >= 1 && v <= kMaxStreams) ? v : kMaxStreams;
  }
  static int64_t xkey(const PTerm& t) { return t.key | ((int64_t)t.in_alt << 56); }   // identity of an input tile stream
  static void simulate_bundles(const std::vector<PRow>& rows, int NX, int Bmax, bool accflag, BundleSim& o) {
    o = BundleSim{};
    size_t ri = 0;
    while (ri < rows.size()) {
      BundleRec B{};
      const size_t ri0 = ri;
      B.len = rows[ri].len;
      B.gate = rows[ri].wait | (rows[ri].sig << 4);
      B.group_begin = (int32_t)o.groups.size();
      std::map<int64_t, int> where;                      // input key -> latest allocation
      std::vector<std::pair<int, int>> last_use;         // per allocation: (group index, term index)
      int next_alloc = 0, nrows = 0;
      bool all_early = true;
      auto resident = [&](int64_t key) {
        auto it = where.find(key);
        return (it != where.end() && next_alloc - it->second <= NX) ? it->second : -1;
      };
      while (ri < rows.size() && nrows < Bmax && rows[ri].len == B.len && next_alloc < 60000 &&
             (rows[ri].wait | (rows[ri].sig << 4)) == B.gate && (nrows == 0 || rows[ri].early == all_early)) {
        const PRow& row = rows[ri];
        all_early = all_early && row.early;
        // A row joins the bundle when it shares an input tile with it.  Rows WITHOUT terms (zero-filled
        // output blocks, e.g. the halo columns of a rank-local adjoint) ride along with whatever bundle
        // is open instead of becoming one-row bundles of their own: those tripled the unit count of the
        // adjoint of a halo-extended block-tridiagonal operator and forced a coarse claim size on it
        // (measured at config 5: 8.7 ms -> 7.7 ms).
        if (nrows > 0 && next_alloc > 0 && !row.terms.empty()) {
          bool share = false;
          for (const PTerm& t : row.terms) share = share || resident(xkey(t)) >= 0;
          if (!share) break;
        }
        BGroupRec cur{};
        int nst = 0;
        bool first = true;
        auto flush = [&](bool row_last) {
          cur.out_off = row.out_off;
          cur.row_in_bundle = nrows;
          cur.flags = (first ? BG_ROW_FIRST : 0) | (row_last ? BG_ROW_LAST : 0) | (accflag ? BG_ACC : 0) |
                      (row.out_alt << BG_OUT_ALT_SHIFT);
          o.sstreams = std::max(o.sstreams, (int)cur.nsstreams);
          o.groups.push_back(cur);
          cur = BGroupRec{};
          nst = 0;
          first = false;
        };
        for (const PTerm& t : row.terms) {
          if (cur.nterms > 0 && (cur.nterms == kGroupTerms || cur.nsstreams + (int)t.sptr.size() > max_group_streams() ||
                                 nst + (int)t.stages.size() > kGroupStages))
            flush(false);
          int gi = (int)o.groups.size();
          int a = resident(xkey(t));
          uint8_t xf = 0;
          if (a >= 0) {
            o.maxdist = std::max(o.maxdist, next_alloc - a);
          } else {
            a = next_alloc;
            const int occ = a - NX;                      // the allocation this one recycles
            if (occ >= 0 && last_use[occ].first >= gi) { // still in use by the group being packed
              flush(false);
              gi = (int)o.groups.size();
            }
            ++next_alloc;
            where[xkey(t)] = a;
            last_use.emplace_back(gi, 0);
            xf = XF_LOAD;
          }
          const int ti = cur.nterms++;
          last_use[a] = {gi, ti};
          BTerm& bt = cur.terms[ti];
          bt.stage0 = (uint8_t)nst;
          bt.nstages = (uint8_t)t.stages.size();
          bt.sstream0 = (uint8_t)cur.nsstreams;
          bt.pattern = (uint8_t)t.pattern;
          bt.sign = (int8_t)t.sign;
          bt.xflags = xf;
          bt.xrel = (uint16_t)a;
          cur.xptr[ti] = t.key;
          cur.xrel_mask |= (1 << ti) | (t.in_alt << (kXAltShift + 2 * ti));
          for (int64_t p : t.sptr) cur.sptr[cur.nsstreams++] = p;
          for (const CStage& c : t.stages) cur.stages[nst++] = c;
        }
        flush(true);
        ++nrows;
        ++ri;
      }
      for (auto& lu : last_use) o.groups[lu.first].terms[lu.second].xflags |= XF_RELEASE;
      B.ngroups = (int32_t)o.groups.size() - B.group_begin;
      B.nrows = nrows;
      B.nx = next_alloc;
      o.max_rows = std::max(o.max_rows, nrows);
      o.bundles.push_back(B);
      o.early.push_back(all_early && nrows > 0);
      o.row_range.push_back({(int)ri0, (int)ri});
    }
  }

  // Largest number of lanes that may issue groups of one unit in parallel such that a lane never
  // waits for an x buffer whose release depends on a group issued by its own batch.
  static int safe_lanes(const BundleSim& sim, int NX, int NS) {
    int G = std::min(NS, 32);
    for (; G > 1; --G) {
      bool ok = true;
      for (const BundleRec& B : sim.bundles) {
        std::vector<int> first_g(B.nx, 0), last_g(B.nx, 0);
        for (int g = 0; g < B.ngroups; ++g) {
          const BGroupRec& r = sim.groups[B.group_begin + g];
          for (int t = 0; t < r.nterms; ++t) {
            if (r.terms[t].xflags & XF_LOAD) first_g[r.terms[t].xrel] = g;
            if (r.terms[t].xflags & XF_RELEASE) last_g[r.terms[t].xrel] = g;
          }
        }
        for (int a = NX; a < B.nx && ok; ++a) ok = last_g[a - NX] / G < first_g[a] / G;
        if (!ok) break;
      }
      if (ok) break;
    }
    return G;
  }

  bool emit_bundle(Entries& es, const Space& out_sp, const Space& in_sp, Ref dst, Ref src, int acc, int hl, int hr,
                   int variant_hint) {
    const size_t esz = dsize(dtype);
    const auto oo = offsets_of(out_sp), io = offsets_of(in_sp);
    std::vector<PRow> rows;
    size_t k = 0;
    for (size_t r = 0; r < out_sp.len.size(); ++r) {
      PRow row;
      row.out_off = dst.off + oo[r];
      row.len = out_sp.len[r];
      while (k < es.size() && es[k].r == (int)r) {
        const Entry& e = es[k++];
        JETS_CHECK(in_sp.len[e.c] == out_sp.len[r], JETS_ERR_SHAPE, "elementwise block (%d,%d) maps %lld -> %lld elements",
                   (int)r, e.c, (long long)in_sp.len[e.c], (long long)out_sp.len[r]);
        row.terms.push_back(make_term(e, acc, (src.off + io[e.c]) * (int64_t)esz, 0));
      }
      if (row.len == 0) continue;
      if (row.terms.empty() && acc != ACC_SET) continue;  // nothing to add
      rows.push_back(std::move(row));
    }
    Step st;
    st.kind = ST_FUSED;
    st.src = src; st.dst = dst; st.acc = acc;
    if (!emit_bundle_rows(rows, st, hl, hr, acc != ACC_SET, variant_hint)) return false;
    st.fused.covers_out = dst.off == 0;   // every non-empty output row is written by this launch (ACC_SET keeps term-less rows)
    plan.steps.push_back(std::move(st));
    return true;
  }

  static PTerm make_term(const Entry& e, int acc, int64_t key, int in_alt) {
    PTerm t;
    t.pattern = classify(e.chain);
    t.sign = (acc == ACC_SUB) ? -e.sign : e.sign;
    t.key = key;
    t.in_alt = in_alt;
    for (const FStage& s : e.chain) {
      CStage cs{};
      cs.op = (uint8_t)s.op; cs.fn = (uint8_t)s.fn; cs.has_stream = s.ptr != nullptr; cs.c0 = s.c0;
      t.stages.push_back(cs);
      if (s.ptr) t.sptr.push_back((int64_t)reinterpret_cast<uintptr_t>(s.ptr));
    }
    return t;
  }

  // Rows (in launch order: units are enumerated bundle-major, so rows listed first are claimed first) ->
  // ONE bundle launch in `st`.  Returns false when the rings do not fit.
  bool emit_bundle_rows(std::vector<PRow>& rows, Step& st, int hl, int hr, bool accflag, int variant_hint) {
    const size_t esz = dsize(dtype);
    DevFused& f = st.fused;
    f.hl = hl; f.hr = hr;
    f.fast = f.use_tma = f.bundle = true;
    if (rows.empty()) {
      plan.engines |= 1;
      return true;
    }
    {   // what the launch writes (relative to `out`) and where its operator state lives
      bool first_o = true;
      for (const PRow& r : rows) {
        if (r.out_alt == 0) {
          const int64_t lo = r.out_off * (int64_t)esz, hi = (r.out_off + r.len) * (int64_t)esz;
          if (first_o) { f.out_lo = lo; f.out_hi = hi; first_o = false; }
          else { f.out_lo = std::min(f.out_lo, lo); f.out_hi = std::max(f.out_hi, hi); }
        }
        for (const PTerm& t : r.terms)
          for (int64_t p : t.sptr) {
            const uintptr_t lo = (uintptr_t)p, hi = lo + (uintptr_t)r.len * esz;
            if (!f.state_hi) { f.state_lo = lo; f.state_hi = hi; }
            else { f.state_lo = std::min(f.state_lo, lo); f.state_hi = std::max(f.state_hi, hi); }
          }
      }
    }
    // reuse distance with an unbounded ring -> how many buffers sharing needs
    BundleSim sim;
    simulate_bundles(rows, kMaxRing, 1 << 30, accflag, sim);
    const int live = std::max(1, std::min(sim.maxdist, 8));
    const int budget = bundle_smem_budget();
    int variant = -1, NX = 0, NS = 0;
    for (int v : {variant_hint, 0}) {
      const int buf = bundle_buf_bytes(v);
      const int slot = sim.sstreams * buf;
      const int D = (budget - live * buf) / (buf + slot);
      int nx = std::min(kMaxRing, std::max(4, live + D));
      if (ctx().bundle_nx > 0) nx = std::min(kMaxRing, ctx().bundle_nx);     // tuning / test overrides
      int ns = slot > 0 ? std::min(kMaxRing, (budget - nx * buf) / slot) : kMaxRing;
      if (ctx().bundle_ns > 0) ns = std::min(ns, ctx().bundle_ns);
      // 16 KB tiles need a state ring at least 3 deep to keep loads in flight while a slot is being
      // consumed (measured on the 4x4 diagonal shape: 2 slots of 4 streams lose 6% to 8 KB tiles)
      const int need = ctx().bundle_ns > 0 ? 1 : (v >= 2 ? 3 : 2);
      if (ns < need || nx * buf + ns * slot > budget) continue;
      variant = v; NX = nx; NS = ns;
      break;
    }
    if (variant < 0) return false;
    // bundle length: long bundles share the most, but there must be enough units to fill the GPU
    const int grid = std::max(1, ctx().sm_count);
    const int64_t cap = (bundle_buf_bytes(variant) - 32) / (int64_t)esz;
    auto units = [&](const BundleSim& s, int64_t te) {
      int64_t u = 0;
      for (const BundleRec& b : s.bundles) u += (b.len + te - 1) / te;
      return u;
    };
    int Bmax = ctx().bundle_bmax > 0 ? ctx().bundle_bmax : (1 << 30);
    while (true) {
      simulate_bundles(rows, NX, Bmax, accflag, sim);
      if (units(sim, cap) >= 2 * (int64_t)grid || sim.max_rows <= 1) break;
      Bmax = std::max(1, std::min(Bmax, sim.max_rows) / 2);
    }
    // Tile length.  With at least four units per CTA the full buffer: every unit costs a pipeline bubble (~2.6 us per
    // CTA on config 1: the rings hold little more than one unit, so the next unit's loads wait for this unit's slots) and
    // full buffers keep the most bytes in flight -- measured on config 1, 848-element tiles (8 whole waves) 74.9 us per
    // pair, 1024-element ones (6.6 waves, ragged) 68.3 us; the dynamic scheduler evens the ragged last wave out.  With
    // fewer units per CTA a ragged wave is a large part of the launch: the largest tile (within 25 % of the capacity)
    // whose unit count fills whole waves.
    int64_t te = cap;
    if (ctx().static_sched || units(sim, cap) < 4 * (int64_t)grid) {
      const int64_t step = 128 / (int64_t)esz;
      double best = -1;
      for (int64_t c = cap; c >= cap - cap / 4 && c >= step; c -= step) {
        const int64_t u = units(sim, c);
        const int64_t waves = (u + grid - 1) / grid;
        const double eff = (double)u / (double)(waves * grid);
        if (eff > best + 1e-3) { best = eff; te = c; }
      }
    }
    if (ctx().tile_elems > 0) te = std::min<int64_t>(cap, std::max<int64_t>(128 / (int64_t)esz, ctx().tile_elems / (128 / (int64_t)esz) * (128 / (int64_t)esz)));
    for (BundleRec& b : sim.bundles) { b.pos0 = 0; b.npos = (int32_t)((b.len + te - 1) / te); }
    // Fine-grained tail.  A unit of a long bundle is a lot of work (config 5: 256 rows x 3 tiles = 12 MB, 280 us), the
    // SMs do not run at the same speed, and the launch ends when the slowest CTA finishes its last unit: traced on
    // config 5, the CTAs finish up to 330 us apart and idle 146 us on average (1.8 % of the launch).  So the LAST tile
    // positions of every long bundle are enumerated again as sub-bundles of an eighth of its rows (an extra 2 input
    // tiles per cut and position: negligible), claimed after all full-height units: the dynamic scheduler evens the
    // finishing times out with them.
    int G_tail = 1 << 30;
    size_t tail_first = (size_t)-1, tail_last = 0;     // [tail_first, tail_last): the tail sub-bundles in sim.bundles
    if (!ctx().no_tail_split && !ctx().static_sched && ctx().bundle_bmax <= 0) {
      int64_t heavy_units = 0;
      int big_rows = 0;
      const int min_rows = ctx().tail_min_rows;
      for (size_t i = 0; i < sim.bundles.size(); ++i)
        if (!sim.early[i] && sim.bundles[i].nrows >= min_rows) { heavy_units += sim.bundles[i].npos; big_rows = std::max(big_rows, sim.bundles[i].nrows); }
      const int64_t min_units = ctx().tail_min_units > 0 ? ctx().tail_min_units : 8 * (int64_t)grid;
      if (heavy_units >= min_units) {
        const int sub_rows = std::max(ctx().tail_sub_min, (big_rows + ctx().tail_div - 1) / ctx().tail_div);
        BundleSim sub;
        simulate_bundles(rows, NX, sub_rows, accflag, sub);
        G_tail = safe_lanes(sub, NX, NS);
        // tail positions: about 1.5 full-height units per CTA in total, shared out over the long bundles by length
        std::vector<BundleRec> mains, tails, rest_late;
        std::vector<char> e_mains, e_tails, e_late;
        std::vector<BGroupRec> groups = sim.groups;
        const int32_t sub_group_base = (int32_t)groups.size();
        groups.insert(groups.end(), sub.groups.begin(), sub.groups.end());
        for (size_t i = 0; i < sim.bundles.size(); ++i) {
          BundleRec b = sim.bundles[i];
          const bool late = (b.gate & ((1 << GF_LO_READY) | (1 << GF_HI_READY))) && b.nrows < min_rows;
          if (sim.early[i] || b.nrows < min_rows) {
            (late ? rest_late : mains).push_back(b);
            (late ? e_late : e_mains).push_back(sim.early[i]);
            continue;
          }
          int32_t pt = (int32_t)std::min<int64_t>(b.npos / 3, (3 * (int64_t)grid * b.npos + 2 * heavy_units - 1) / (2 * heavy_units));
          if (pt < 1) { mains.push_back(b); e_mains.push_back(0); continue; }
          BundleRec m = b;
          m.npos = b.npos - pt;
          mains.push_back(m); e_mains.push_back(0);
          for (size_t j = 0; j < sub.bundles.size(); ++j) {       // the sub-bundles made of this bundle's rows
            if (sub.row_range[j].first < sim.row_range[i].first || sub.row_range[j].second > sim.row_range[i].second) continue;
            BundleRec t = sub.bundles[j];
            t.group_begin += sub_group_base;
            t.pos0 = b.npos - pt;
            t.npos = pt;
            tails.push_back(t); e_tails.push_back(0);
          }
        }
        sim.groups.swap(groups);
        sim.bundles.clear(); sim.early.clear();
        tail_first = mains.size(); tail_last = mains.size() + tails.size();
        for (auto* v : {&mains, &tails, &rest_late}) sim.bundles.insert(sim.bundles.end(), v->begin(), v->end());
        for (auto* v : {&e_mains, &e_tails, &e_late}) sim.early.insert(sim.early.end(), v->begin(), v->end());
      }
    }
    for (BGroupRec& gr : sim.groups)          // split every allocation index for the kernel (no division per term)
      for (int t = 0; t < gr.nterms; ++t) {
        BTerm& bt = gr.terms[t];
        bt.xflags = (uint8_t)((bt.xflags & 3) | (((bt.xrel / NX) & 1) << 2) | ((bt.xrel % NX) << 4));
      }
    int64_t unit = 0, rowtiles = 0;
    for (BundleRec& b : sim.bundles) {
      b.unit_begin = unit;
      b.pbase = rowtiles;
      unit += b.npos;
      rowtiles += (int64_t)b.npos * b.nrows;
    }
    f.nrowtiles = rowtiles;
    { static const int cw[6] = {16, 8, 16, 8, 30, 24}; f.consumer_warps = cw[variant >= 0 && variant < 6 ? variant : 0]; }
    f.variant = variant;
    f.NX = NX; f.NS = NS; f.sstreams = sim.sstreams;
    f.G = std::min(safe_lanes(sim, NX, NS), G_tail);
    f.tile_elems = (int)te;
    f.nbundles = (int32_t)sim.bundles.size();
    f.nunits = unit;
    f.nrows = (int32_t)rows.size();
    f.ntiles = unit;
    for (const BundleRec& b : sim.bundles) {        // units behind every cross-rank signal
      const int64_t u = b.npos;
      for (int s = 0; s < kGateFlags; ++s)
        if ((b.gate >> (4 + s)) & 1) f.sig_total[s] += (int32_t)u;
      if (b.gate) f.gated = true;
    }
    // Dynamic claims: sized per bundle for roughly equal work (a unit's work ~ its number of term groups): the
    // heaviest bundle is claimed one unit at a time, light ones (a push row, a term-less row) many at a time so
    // that the claim atomics stay off their critical path -- but never so coarse that a CTA gets fewer than ~8
    // claims (a coarse claim on a heavy bundle is a long tail: measured 14% on the gated config-5 launch).
    int max_groups = 1;
    for (const BundleRec& b : sim.bundles) max_groups = std::max(max_groups, b.ngroups);
    int64_t nclaims = 0;
    for (int shrink = 0; ; ++shrink) {
      nclaims = 0;
      bool any = false;
      for (size_t bi = 0; bi < sim.bundles.size(); ++bi) {
        BundleRec& b = sim.bundles[bi];
        const int64_t u = b.npos;
        // equal work per claim ...
        int64_t ch = std::max<int64_t>(1, max_groups / std::max(1, b.ngroups));
        // ... and at least what the producer issues side by side for a short bundle (several units per batch)
        if (2 * b.ngroups <= f.G) {
          int64_t side = f.G / b.ngroups;
          if (b.nx > 0) side = std::min<int64_t>(side, NX / b.nx);
          ch = std::max(ch, side);
        }
        // bundles that wait for a neighbour's "ready" flag run at the very END of the launch: small claims there, so
        // that they even out the CTAs' finishing times (N=8 trace: 30-unit claims = 1.7 claims per CTA left the CTAs
        // of the adjoint finishing up to 100 us apart, and the skew came back as start waits in the next forward)
        if ((b.gate & ((1 << GF_LO_READY) | (1 << GF_HI_READY))) && b.ngroups * 4 <= max_groups) ch = std::min<int64_t>(ch, 4);
        if (bi >= tail_first && bi < tail_last) ch = 1;       // the tail exists for its granularity
        ch = std::max<int64_t>(1, std::min<int64_t>(32, ch) >> shrink);
        any = any || ch > 1;
        b.chunk = (int32_t)ch;
        b.claim_begin = (int32_t)nclaims;
        b.nclaims = (int32_t)((u + ch - 1) / ch);
        nclaims += b.nclaims;
      }
      if (nclaims >= 8 * (int64_t)grid || !any) break;    // keep at least ~8 claims per CTA: halve the claims and retry
    }
    f.nclaims = nclaims;
    const int chunk = 1;
    f.chunk = 1;
    {   // leading bundles made of early rows only
      bool lead = true;
      for (size_t i = 0; i < sim.bundles.size(); ++i) {
        const BundleRec& b = sim.bundles[i];
        lead = lead && sim.early[i];
        if (lead) f.early_claims += (b.npos + b.chunk - 1) / b.chunk;
      }
    }
    const size_t gb = (sim.groups.size() * sizeof(BGroupRec) + 255) & ~(size_t)255;
    const size_t bb = (sim.bundles.size() * sizeof(BundleRec) + 255) & ~(size_t)255;
    std::vector<char> host(gb + bb + 256, 0);     // tables, then the (zeroed) scheduler counters
    memcpy(host.data(), sim.groups.data(), sim.groups.size() * sizeof(BGroupRec));
    memcpy(host.data() + gb, sim.bundles.data(), sim.bundles.size() * sizeof(BundleRec));
    char* blob = nullptr;
    CUDA_TRY(cudaMalloc(&blob, host.size()));
    CUDA_TRY(cudaMemcpy(blob, host.data(), host.size(), cudaMemcpyHostToDevice));
    plan.blobs.push_back(blob);
    f.blob = blob;
    f.bgroups = reinterpret_cast<BGroupRec*>(blob);
    f.bundles = reinterpret_cast<BundleRec*>(blob + gb);
    f.table_bytes = gb + bb;
    f.sched = (!ctx().static_sched && nclaims > (int64_t)grid) ? reinterpret_cast<int32_t*>(blob + gb + bb) : nullptr;
    f.sig_done = reinterpret_cast<int32_t*>(blob + gb + bb + 64);
    plan.engines |= 1 | 32;
    if (getenv("JETS_B200_PLAN_DEBUG")) {
      int64_t ng = 0, nxsum = 0, maxg = 0;
      for (const BundleRec& b : sim.bundles) { ng += b.ngroups; nxsum += b.nx; maxg = std::max<int64_t>(maxg, b.ngroups); }
      fprintf(stderr, "[jets plan] bundle engine: rows=%zu bundles=%zu groups=%lld (max %lld per bundle) x-allocs=%lld units=%lld tile=%lld elems "
              "variant=%d NX=%d NS=%d sstreams=%d G=%d chunk=%d dyn=%d\n", rows.size(), sim.bundles.size(), (long long)ng, (long long)maxg,
              (long long)nxsum, (long long)unit, (long long)te, variant, NX, NS, sim.sstreams, f.G, chunk, f.sched != nullptr);
      if (atoi(getenv("JETS_B200_PLAN_DEBUG")) >= 2)
        for (size_t i = 0; i < sim.bundles.size(); ++i) {
          const BundleRec& b = sim.bundles[i];
          fprintf(stderr, "[jets plan]   bundle %zu: rows=%d groups=%d x-allocs=%d positions=%d (from %d) claim=%d units gate=0x%x early=%d\n", i, b.nrows,
                  b.ngroups, b.nx, b.npos, b.pos0, b.chunk, (unsigned)b.gate, (int)sim.early[i]);
        }
    }
    return true;
  }


  // ---- distributed block-banded apply (dist.cu) ------------------------------------------------
  // A_loc is this rank's nloc x (nloc + 2h) block rows over the halo-extended domain
  // [h blocks of the previous rank | nloc own blocks | h blocks of the next rank].  The launch works on the
  // rank's OWN shards (`in`/`out` hold nloc blocks); halo data lives in the exchange arena:
  //   forward   push rows copy the first / last h own blocks of `in` into the neighbours' halo buffers
  //             (peer stores) and raise their "ready" flag; rows that read a halo buffer wait for this
  //             rank's "ready" flag, are enumerated last, and report "done" back to the writer.
  //   adjoint   the partial sums for the neighbours' columns (src/Jets.jl:1039-1055 restricted to this rank's
  //             rows) are stored straight into the neighbours' staging buffers; the owner adds them as one
  //             more term of its row sum -- the previous rank's partial FIRST, the next rank's LAST, which
  //             is the single-GPU left-to-right order (:1049), so a block-tridiagonal result is bit-identical.
  void emit_banded(jets_op A, int h, const BandedSel& sel) {
    const bool adj = sel.mode == JETS_MODE_DFT;
    const int nloc = A->R;
    const size_t esz = dsize(dtype);
    Entries es;
    Space isp, osp;
    JETS_CHECK(expand(A, sel.mode, es, isp, osp), JETS_ERR_UNSUPPORTED,
               "distributed banded apply: the local block rows must be elementwise/stencil operators (one fused launch)");
    int hl, hr;
    halo_of(es, hl, hr);
    JETS_CHECK(hl <= 1 && hr <= 1, JETS_ERR_UNSUPPORTED, "distributed banded apply: stencil halo too wide for fusion");
    std::stable_sort(es.begin(), es.end(), [](const Entry& x, const Entry& y) { return x.r != y.r ? x.r < y.r : x.c < y.c; });
    for (const Entry& e : es) {
      JETS_CHECK(classify(e.chain) != PAT_GENERIC, JETS_ERR_UNSUPPORTED, "distributed banded apply: block (%d,%d) has no straight-line kernel path", e.r, e.c);
      for (const FStage& s : e.chain)
        JETS_CHECK(stream_tma_ok(s), JETS_ERR_UNSUPPORTED, "distributed banded apply: operator state must live in 16-byte aligned library-owned buffers");
    }
    const Space& dom = A->dom;   // nloc + 2h blocks
    const Space& rng = A->rng;   // nloc blocks
    for (auto l : dom.len) JETS_CHECK((l * esz) % 16 == 0, JETS_ERR_UNSUPPORTED, "distributed banded apply: block lengths must be multiples of 16 bytes");
    for (auto l : rng.len) JETS_CHECK((l * esz) % 16 == 0, JETS_ERR_UNSUPPORTED, "distributed banded apply: block lengths must be multiples of 16 bytes");
    // element offsets inside: the own-domain shard, the range shard, the lo / hi halo (or staging) buffers
    std::vector<int64_t> own_off(nloc + 1, 0), rng_off(nloc + 1, 0), lo_off(h + 1, 0), hi_off(h + 1, 0), slo_off(h + 1, 0), shi_off(h + 1, 0);
    for (int b = 0; b < nloc; ++b) { own_off[b + 1] = own_off[b] + dom.len[h + b]; rng_off[b + 1] = rng_off[b] + rng.len[b]; }
    for (int k = 0; k < h; ++k) {
      lo_off[k + 1] = lo_off[k] + dom.len[k];                 // forward halo / partials for the previous rank
      hi_off[k + 1] = hi_off[k] + dom.len[nloc + h + k];
      slo_off[k + 1] = slo_off[k] + dom.len[h + k];           // adjoint staging: partials received for my first / last h columns
      shi_off[k + 1] = shi_off[k] + dom.len[nloc + k];
    }
    std::vector<PRow> early, plain, late;
    auto copy_term = [&](int64_t key_bytes, int in_alt) {
      Entry e;   // empty chain = PAT_COPY
      return make_term(e, ACC_SET, key_bytes, in_alt);
    };
    if (!adj) {
      if (sel.send_prev && sel.has_prev && !sel.pull)
        for (int b = 0; b < h; ++b) {       // my first h own blocks are the previous rank's hi halo
          PRow row;
          row.early = true;
          row.out_alt = 1; row.out_off = slo_off[b]; row.len = dom.len[h + b];
          row.wait = 1 << GF_PREV_DONE; row.sig = 1 << GS_PREV_HI_READY;
          row.terms.push_back(copy_term(own_off[b] * (int64_t)esz, 0));
          if (row.len) early.push_back(std::move(row));
        }
      if (sel.send_next && sel.has_next && !sel.pull)
        for (int k = 0; k < h; ++k) {       // my last h own blocks are the next rank's lo halo
          const int b = nloc - h + k;
          PRow row;
          row.early = true;
          row.out_alt = 2; row.out_off = shi_off[k]; row.len = dom.len[h + b];
          row.wait = 1 << GF_NEXT_DONE; row.sig = 1 << GS_NEXT_LO_READY;
          row.terms.push_back(copy_term(own_off[b] * (int64_t)esz, 0));
          if (row.len) early.push_back(std::move(row));
        }
      size_t k = 0;
      for (int r = 0; r < nloc; ++r) {
        PRow row;
        row.out_off = rng_off[r]; row.len = rng.len[r];
        while (k < es.size() && es[k].r < r) ++k;
        const bool want = r >= sel.row_begin && r < sel.row_end;
        while (k < es.size() && es[k].r == r) {
          const Entry& e = es[k++];
          if (!want) continue;
          JETS_CHECK(dom.len[e.c] == rng.len[r], JETS_ERR_SHAPE, "elementwise block (%d,%d) maps %lld -> %lld elements", r, e.c,
                     (long long)dom.len[e.c], (long long)rng.len[r]);
          if (e.c < h) {
            JETS_CHECK(sel.has_prev, JETS_ERR_INVALID, "distributed banded apply: block (%d,%d) reads the previous rank's halo but this is the first rank", r, e.c);
            row.terms.push_back(make_term(e, ACC_SET, lo_off[e.c] * (int64_t)esz, 1));
            row.wait |= 1 << GF_LO_READY; row.sig |= 1 << GS_PREV_NEXT_DONE;
          } else if (e.c >= nloc + h) {
            JETS_CHECK(sel.has_next, JETS_ERR_INVALID, "distributed banded apply: block (%d,%d) reads the next rank's halo but this is the last rank", r, e.c);
            row.terms.push_back(make_term(e, ACC_SET, hi_off[e.c - nloc - h] * (int64_t)esz, 2));
            row.wait |= 1 << GF_HI_READY; row.sig |= 1 << GS_NEXT_PREV_DONE;
          } else {
            row.terms.push_back(make_term(e, ACC_SET, own_off[e.c - h] * (int64_t)esz, 0));
          }
        }
        if (!want || row.len == 0) continue;
        // pull mode: the halo terms read the neighbour's vector in place (final since its launch began), so the
        // rows stay in the main bundle -- no cut of the sliding input window, no halo copy
        ((row.wait && !sel.pull) ? late : plain).push_back(std::move(row));
      }
      if (sel.pull) {
        // one bundle: every unit waits for (and reports to) the same flags
        int w = 0, sg = 0;
        for (auto& r : plain) { w |= r.wait; sg |= r.sig; }
        for (auto& r : plain) { r.wait = w; r.sig = sg; }
      }
    } else {
      // entries: r = extended column (output), c = local row (input block of d)
      size_t k = 0;
      for (int j = 0; j < nloc + 2 * h; ++j) {
        PRow row;
        row.len = dom.len[j];
        bool want;
        if (j < h) {                               // partial sums for the previous rank's last h columns
          want = sel.send_prev && sel.has_prev;
          row.out_alt = 1; row.out_off = lo_off[j];
          row.wait = 1 << GF_PREV_DONE; row.sig = 1 << GS_PREV_HI_READY;
        } else if (j >= nloc + h) {                // ... for the next rank's first h columns
          want = sel.send_next && sel.has_next;
          row.out_alt = 2; row.out_off = hi_off[j - nloc - h];
          row.wait = 1 << GF_NEXT_DONE; row.sig = 1 << GS_NEXT_LO_READY;
        } else {
          const int b = j - h;
          want = b >= sel.row_begin && b < sel.row_end;
          row.out_off = own_off[b];
          if (want && sel.has_prev && b < h) {     // the previous rank's partial enters first (:1049 order)
            row.terms.push_back(copy_term(slo_off[b] * (int64_t)esz, 1));
            row.wait |= 1 << GF_LO_READY; row.sig |= 1 << GS_PREV_NEXT_DONE;
          }
        }
        while (k < es.size() && es[k].r < j) ++k;
        while (k < es.size() && es[k].r == j) {
          const Entry& e = es[k++];
          if (!want) continue;
          JETS_CHECK(rng.len[e.c] == dom.len[j], JETS_ERR_SHAPE, "elementwise block (%d,%d) maps %lld -> %lld elements", e.c, j,
                     (long long)dom.len[j], (long long)rng.len[e.c]);
          row.terms.push_back(make_term(e, ACC_SET, rng_off[e.c] * (int64_t)esz, 0));
        }
        if (want && j >= h && j < nloc + h && sel.has_next && j - h >= nloc - h) {   // the next rank's partial enters last
          row.terms.push_back(copy_term(shi_off[j - h - (nloc - h)] * (int64_t)esz, 2));
          row.wait |= 1 << GF_HI_READY; row.sig |= 1 << GS_NEXT_PREV_DONE;
        }
        if (!want || row.len == 0) continue;
        if (j < h || j >= nloc + h) { row.early = true; early.push_back(std::move(row)); }
        else (row.wait ? late : plain).push_back(std::move(row));
      }
      // (Measured: letting the partial-sum rows ride in the main bundle -- to share its d tiles -- turns the
      // neighbours' "partial ready" signal into a barrier in the middle of the launch, the columns that add a
      // partial cannot start before EVERY main unit is done: 1.047 ms against 1.025 ms at 32 local rows.  They
      // stay in the early queue.)
    }
    std::vector<PRow> rows;
    rows.reserve(early.size() + plain.size() + late.size());
    for (auto* v : {&early, &plain, &late})
      for (auto& r : *v) rows.push_back(std::move(r));
    Step st;
    st.kind = ST_FUSED;
    st.src = Ref{0, 0}; st.dst = Ref{1, 0}; st.acc = ACC_SET;
    int64_t longest = 0;
    for (auto& r : rows) longest = std::max(longest, r.len);
    const int variant = (ctx().fast_variant >= 0 && ctx().fast_variant <= 5) ? ctx().fast_variant : ((longest * (int64_t)esz >= 4 * 16384) ? 2 : 0);
    JETS_CHECK(emit_bundle_rows(rows, st, hl, hr, false, variant), JETS_ERR_UNSUPPORTED,
               "distributed banded apply: the operator's rows do not fit the bundle kernel's shared-memory rings");
    st.fused.sig_owned = sel.owned;
    st.fused.gated = sel.has_prev || sel.has_next;   // launched with the arena's flag words and halo bases
    plan.steps.push_back(std::move(st));
  }

  // Dense entries sharing one orientation; groups = distinct output blocks.
  void emit_gemv(std::vector<DBlock>& blocks, int acc) {
    if (blocks.empty()) return;
    std::stable_sort(blocks.begin(), blocks.end(),
                     [](const DBlock& x, const DBlock& y) { return x.out_off < y.out_off; });
    Step st;
    st.kind = ST_GEMV;
    st.acc = acc;
    std::vector<int32_t> row_ptr{0}, tile_ptr{0};
    const bool trans = blocks[0].trans != 0;
    for (size_t i = 0; i < blocks.size(); ++i) {
      const bool last = i + 1 == blocks.size() || blocks[i + 1].out_off != blocks[i].out_off;
      if (last) {
        row_ptr.push_back((int32_t)i + 1);
        int32_t nt;
        gemv_tile_count(dtype, trans, trans ? blocks[i].cols : blocks[i].rows, &nt);
        tile_ptr.push_back(tile_ptr.back() + nt);
      }
    }
    st.n_out_rows = (int32_t)row_ptr.size() - 1;
    st.gemv_tiles = tile_ptr.back();
    st.dblocks = blocks;
    const size_t nb = blocks.size() * sizeof(DBlock);
    const size_t nb_al = (nb + 255) & ~(size_t)255;
    std::vector<char> host(nb_al + 2 * row_ptr.size() * sizeof(int32_t));
    memcpy(host.data(), blocks.data(), nb);
    memcpy(host.data() + nb_al, row_ptr.data(), row_ptr.size() * sizeof(int32_t));
    memcpy(host.data() + nb_al + row_ptr.size() * sizeof(int32_t), tile_ptr.data(), tile_ptr.size() * sizeof(int32_t));
    char* blob = nullptr;
    CUDA_TRY(cudaMalloc(&blob, host.size()));
    CUDA_TRY(cudaMemcpy(blob, host.data(), host.size(), cudaMemcpyHostToDevice));
    plan.blobs.push_back(blob);
    st.d_dblocks = reinterpret_cast<DBlock*>(blob);
    st.d_row_ptr = reinterpret_cast<int32_t*>(blob + nb_al);
    if (!trans && !is_cplx(dtype) && st.gemv_tiles > 0 && st.gemv_tiles < 2 * (int64_t)ctx().sm_count) {
      // a block-row shard of a wide operator has few row tiles: split every tile's blocks over several CTAs
      int min_entries = 1 << 30;
      for (size_t g = 0; g + 1 < row_ptr.size(); ++g) min_entries = std::min(min_entries, row_ptr[g + 1] - row_ptr[g]);
      const int64_t want = (4 * (int64_t)ctx().sm_count + st.gemv_tiles - 1) / st.gemv_tiles;
      const int ks = (int)std::max<int64_t>(1, std::min<int64_t>(want, min_entries));
      if (ks > 1) {
        const int tm = 32 * (dtype == JETS_F32 ? 4 : 2);
        double* part = nullptr;
        CUDA_TRY(cudaMalloc(&part, (size_t)ks * st.gemv_tiles * tm * sizeof(double)));
        plan.blobs.push_back(part);
        st.gemv_ksplit = ks;
        st.gemv_partials = part;
      }
    }
    plan.engines |= 4;
    plan.steps.push_back(std::move(st));
  }

  // Is `a` (possibly under adjoint / linear views) a plain dense leaf?  Returns the leaf and the
  // effective orientation for `mode`.
  static jets_op dense_leaf(jets_op a, int mode, bool& trans) {
    while (true) {
      if (a->kind == K_LNVIEW) { mode = map_mode_lnview(mode); a = a->kids[0]; }
      else if (a->kind == K_ADJ) { mode = map_mode_adj(mode); a = a->kids[0]; }
      else break;
    }
    if (a->kind != K_DENSE) return nullptr;
    trans = (mode == JETS_MODE_DFT);
    return a;
  }

  // Dense tables of one apply: single-vector GEMV entries (n: A*x, t: A'*x) and, for leaves with
  // several right-hand sides that qualify for the tensor-core path, block-level entries (tn / tt).
  struct DenseTables {
    std::vector<DBlock> n, t, tn, tt;
    bool empty() const { return n.empty() && t.empty() && tn.empty() && tt.empty(); }
    int kinds() const { return !n.empty() + !t.empty() + !tn.empty() + !tt.empty(); }
  };

  void push_dense(DenseTables& D, jets_op leaf, bool trans, int64_t out_off, int64_t in_off) {
    if (leaf->nrhs >= 2) {
      DBlock b{};
      b.A = leaf->w->ptr();
      b.rows = (int32_t)leaf->rows; b.cols = (int32_t)leaf->cols; b.lda = (int32_t)leaf->rows;
      b.trans = trans; b.nrhs = (int32_t)leaf->nrhs;
      b.in_off = in_off; b.out_off = out_off;
      std::vector<DBlock>& v = trans ? D.tt : D.tn;
      std::vector<DBlock> probe{b};
      if (gemm_tc_eligible(dtype, probe) && (v.empty() || v[0].nrhs == b.nrhs)) {
        v.push_back(b);
        return;
      }
    }
    for (int64_t k = 0; k < leaf->nrhs; ++k) {
      DBlock b{};
      b.A = leaf->w->ptr();
      b.rows = (int32_t)leaf->rows; b.cols = (int32_t)leaf->cols; b.lda = (int32_t)leaf->rows;
      b.trans = trans; b.nrhs = 1;
      b.in_off = in_off + k * (trans ? leaf->rows : leaf->cols);
      b.out_off = out_off + k * (trans ? leaf->cols : leaf->rows);
      (trans ? D.t : D.n).push_back(b);
    }
  }

  // Tensor-core GEMM over a table of block-level dense entries (all one orientation, same nrhs).
  void emit_gemm_tc(std::vector<DBlock>& blocks, int acc, Ref src, Ref dst) {
    if (blocks.empty()) return;
    std::stable_sort(blocks.begin(), blocks.end(),
                     [](const DBlock& x, const DBlock& y) { return x.out_off < y.out_off; });
    Step st;
    st.kind = ST_GEMM_TC;
    st.acc = acc;
    st.src = src; st.dst = dst;
    st.dblocks = blocks;
    gemm_tc_prepare(st, plan);
    plan.engines |= 8;
    plan.steps.push_back(std::move(st));
  }

  // Emits every table of D (first one with `acc`, the rest accumulating on top).
  void emit_dense(DenseTables& D, int acc, Ref src, Ref dst) {
    int cur = acc;
    auto next = [&]() { cur = (acc == ACC_SET) ? ACC_ADD : acc; };
    for (auto* v : {&D.n, &D.t}) {
      if (v->empty()) continue;
      emit_gemv(*v, cur);
      plan.steps.back().src = Ref{src.which, 0};
      plan.steps.back().dst = Ref{dst.which, 0};
      next();
    }
    for (auto* v : {&D.tn, &D.tt}) {
      if (v->empty()) continue;
      emit_gemm_tc(*v, cur, Ref{src.which, 0}, Ref{dst.which, 0});
      next();
    }
  }

  static int combine(int acc, int sgn) {  // accumulate mode for a signed term under `acc`
    const int eff = (acc == ACC_SUB ? -1 : 1) * sgn;
    return eff > 0 ? ACC_ADD : ACC_SUB;
  }

  void lower(jets_op a, int mode, Ref dst, Ref src, int acc) {
    {
      Entries es;
      Space isp, osp;
      int hl = 0, hr = 0;
      if (expand(a, mode, es, isp, osp)) {
        halo_of(es, hl, hr);
        if (hl <= 1 && hr <= 1) {
          emit_fused(es, osp, isp, dst, src, acc);
          return;
        }
      }
    }
    switch (a->kind) {
      case K_LNVIEW: lower(a->kids[0], map_mode_lnview(mode), dst, src, acc); return;
      case K_ADJ: lower(a->kids[0], map_mode_adj(mode), dst, src, acc); return;
      case K_DENSE: {
        DenseTables D;
        push_dense(D, a, mode == JETS_MODE_DFT, dst.off, src.off);
        emit_dense(D, acc, src, dst);
        return;
      }
      case K_RESTRICT: {
        // forward d = m[idx]; adjoint m[idx] = d, zero elsewhere (zero-filled here when the apply overwrites)
        const bool adj = (mode == JETS_MODE_DFT);
        if (adj && acc == ACC_SET) emit_fill0(dst, a->dom.total());
        Step st;
        st.kind = ST_GATHER;
        st.src = src; st.dst = dst;
        st.acc = (adj && acc == ACC_SET) ? ACC_ADD : acc;
        st.g_idx = a->gidx.get(); st.g_idx64 = a->gidx64 ? 1 : 0;
        st.g_n = a->rng.total();
        st.g_scatter = adj ? 1 : 0;
        plan.engines |= 64;
        plan.steps.push_back(std::move(st));
        return;
      }
      case K_COMPOSE: {
        plan.engines |= 16;
        const std::vector<jets_op> seq = app_order(a, mode);
        size_t i = 0;
        Ref cur = src;
        while (i < seq.size()) {
          // longest fusible run starting at i
          size_t best = 0;
          Entries best_es;
          Space best_in, best_out;
          for (size_t j = seq.size(); j > i; --j) {
            std::vector<jets_op> grp(seq.begin() + i, seq.begin() + j);
            Entries es;
            Space gi, go;
            int hl, hr;
            if (expand_seq(grp, mode, es, gi, go)) {
              halo_of(es, hl, hr);
              if (hl <= 1 && hr <= 1) { best = j; best_es.swap(es); best_in = gi; best_out = go; break; }
            }
          }
          const size_t j = best ? best : i + 1;
          const bool last = (j == seq.size());
          const Space& osp = out_space(seq[j - 1], mode);
          Ref nxt = last ? dst : new_tmp(osp.total());
          const int nacc = last ? acc : ACC_SET;
          if (best) emit_fused(best_es, best_out, best_in, nxt, cur, nacc);
          else lower(seq[i], mode, nxt, cur, nacc);
          cur = nxt;
          i = j;
        }
        return;
      }
      case K_SUM: {
        plan.engines |= 16;
        int first = acc;
        if (acc == ACC_SET) {
          if (a->sgn[0] > 0) first = ACC_SET;
          else { emit_fill0(dst, out_space(a, mode).total()); first = ACC_SUB; }
        } else {
          first = combine(acc, a->sgn[0]);
        }
        for (size_t k = 0; k < a->kids.size(); ++k) {
          const int m = (k == 0) ? first : combine(acc == ACC_SET ? ACC_ADD : acc, a->sgn[k]);
          lower(a->kids[k], mode, dst, src, m);
        }
        return;
      }
      case K_BLOCK: {
        const bool adj = (mode == JETS_MODE_DFT);
        const Space& osp = out_space(a, mode);
        const Space& isp = in_space(a, mode);
        const auto oo = offsets_of(osp), io = offsets_of(isp);
        const int nout = (int)osp.len.size();
        DenseTables D;
        Entries fes;
        struct Rest { jets_op op; int o, i; };
        std::vector<Rest> rest;
        std::vector<int> covered(nout, 0);
        for (int c = 0; c < a->C; ++c)
          for (int r = 0; r < a->R; ++r) {
            jets_op kid = a->kids[r + (size_t)c * a->R];
            const int o = adj ? c : r, i = adj ? r : c;
            bool trans = false;
            if (jets_op leaf = dense_leaf(kid, mode, trans)) {
              push_dense(D, leaf, trans, dst.off + oo[o], src.off + io[i]);
              covered[o] |= 1;
              continue;
            }
            Entries es;
            Space ki, ko;
            int hl, hr;
            if (expand(kid, mode, es, ki, ko)) {
              halo_of(es, hl, hr);
              bool ok = hl <= 1 && hr <= 1 && is_flat(ki) && is_flat(ko);
              for (auto& x : es) ok = ok && x.r == 0 && x.c == 0;
              if (ok) {
                for (auto& x : es) { x.r = o; x.c = i; fes.push_back(std::move(x)); }
                continue;
              }
            }
            rest.push_back({kid, o, i});
          }
        // Pure dense, one table, every output block covered: SET directly (one launch).
        const bool pure = fes.empty() && rest.empty() && D.kinds() == 1 &&
                          std::all_of(covered.begin(), covered.end(), [](int v) { return v != 0; });
        int cur = acc;
        if (pure) {
          emit_dense(D, acc, src, dst);
          return;
        }
        plan.engines |= 16;
        // first the fused part (it can SET: rows without fused terms are zero-filled by it) ...
        emit_fused(fes, osp, isp, dst, src, acc);
        cur = (acc == ACC_SET) ? ACC_ADD : acc;
        // ... then dense tables and whatever is left accumulate on top
        emit_dense(D, cur, src, dst);
        for (auto& x : rest) {
          Ref d = dst, s = src;
          d.off += oo[x.o];
          s.off += io[x.i];
          lower(x.op, mode, d, s, cur);
        }
        return;
      }
      default: JETS_FAIL(JETS_ERR_UNSUPPORTED, "internal: cannot lower operator kind %d", (int)a->kind);
    }
  }
};

}  // namespace

Plan::~Plan() {
  for (void* p : tmps) cudaFree(p);
  for (void* p : blobs) cudaFree(p);
  for (jets_buf b : point_bufs) jets_buf_destroy(b);
}
bool Plan::valid() const {
  if (!uses_point) return true;
  if (points.size() > kMaxTrackedPoints) return version == g_epoch;
  for (auto& pr : points)
    if (!pr.first->mo || pr.first->mo->ptr() != pr.second) return false;
  return true;
}

namespace {
// Collects the linearization points a plan reads while it is being built (RAII around the builder calls).
struct PointScope {
  std::vector<std::pair<jets_op, const void*>> pts;
  PointScope() { g_points = &pts; }
  ~PointScope() { g_points = nullptr; }
  void finish(Plan& p) {
    p.uses_point = !pts.empty();
    if (pts.size() <= Plan::kMaxTrackedPoints) {
      std::sort(pts.begin(), pts.end());
      pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
      p.points = std::move(pts);
      for (auto& pr : p.points) { pr.first->mo->refs++; p.point_bufs.push_back(pr.first->mo); }
    } else {
      p.points.assign(Plan::kMaxTrackedPoints + 1, {nullptr, nullptr});   // marker: too many to track -> epoch
    }
    p.version = g_epoch;
  }
};
}  // namespace

std::shared_ptr<Plan> build_plan(jets_op a, int mode, int accumulate, bool io_ok, int engine) {
  auto plan = std::make_shared<Plan>();
  g_esz = dsize(a->dtype);
  g_cplx = is_cplx(a->dtype);
  PointScope scope;
  Builder b{*plan, a->dtype, io_ok, engine};
  int acc = ACC_SET;
  if (accumulate) {
    // quirk Q1 (src/Jets.jl:1001,1024): only a forward block apply with ncol>1 accumulates
    jets_op t = a;
    int m = mode;
    while (t->kind == K_LNVIEW) { m = map_mode_lnview(m); t = t->kids[0]; }
    if (t->kind == K_BLOCK && t->C > 1 && m != JETS_MODE_DFT) acc = ACC_ADD;
  }
  b.lower(a, mode, Ref{1, 0}, Ref{0, 0}, acc);
  scope.finish(*plan);
  return plan;
}

std::shared_ptr<Plan> build_banded_plan(jets_op A_loc, int halo, const BandedSel& sel) {
  auto plan = std::make_shared<Plan>();
  g_esz = dsize(A_loc->dtype);
  g_cplx = is_cplx(A_loc->dtype);
  JETS_CHECK(!g_cplx, JETS_ERR_UNSUPPORTED, "distributed banded apply: complex eltypes are not implemented");
  PointScope scope;
  Builder b{*plan, A_loc->dtype, true, 0};
  b.emit_banded(A_loc, halo, sel);
  scope.finish(*plan);
  return plan;
}

// out = cA*(A in) + cO*out fused into the store epilogue when the whole apply is ONE bundle launch
// that writes every output row (otherwise the caller stages through a temporary).
bool run_plan_axpby(Plan& p, int dtype, char* in, char* out, const ApplyCoef& coef) {
  if (p.steps.size() != 1) return false;
  Step& st = p.steps[0];
  if (st.kind != ST_FUSED || !st.fused.bundle || st.acc != ACC_SET || st.src.which != 0 || st.dst.which != 1 ||
      !st.fused.covers_out || (st.fused.variant != 0 && st.fused.variant != 2))   // the epilogue's tile shapes
    return false;
  launch_fused_bundle(st.fused, dtype, in, out, ctx().stream, &coef);
  return true;
}

void run_plan(Plan& p, int dtype, char* in, char* out) {
  Context& c = ctx();
  auto base = [&](const Ref& r) -> char* {
    if (r.which == 0) return in;
    if (r.which == 1) return out;
    return reinterpret_cast<char*>(p.tmps[r.which - 2]) + kGuardBytes;
  };
  for (Step& st : p.steps) {
    switch (st.kind) {
      case ST_FUSED:
        if (st.fused.bundle) launch_fused_bundle(st.fused, dtype, base(st.src), base(st.dst), c.stream);
        else if (st.fused.fast) launch_fused_fast(st.fused, dtype, base(st.src), base(st.dst), c.stream);
        else launch_fused(st.fused, dtype, base(st.src), base(st.dst), c.stream);
        break;
      case ST_GEMV: launch_gemv(st, dtype, base(st.src), base(st.dst), c.stream); break;
      case ST_GEMM_TC: launch_gemm_tc(st, base(st.src), base(st.dst), c.stream); break;
      case ST_GATHER:
        vec_gather(dtype, base(st.dst) + st.dst.off * dsize(dtype), base(st.src) + st.src.off * dsize(dtype), st.g_idx, st.g_idx64,
                   st.g_n, st.g_scatter, st.acc, c.stream);
        break;
      case ST_FILL0:
        if (is_cplx(dtype)) cvec_fill(dtype, base(st.dst) + st.dst.off * dsize(dtype), st.fill_len, 0.0, 0.0, c.stream);
        else vec_fill(dtype, base(st.dst) + st.dst.off * dsize(dtype), st.fill_len, 0.0, c.stream);
        break;
    }
  }
}

}  // namespace jets
