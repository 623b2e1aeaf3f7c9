// Bundle engine: the fused block apply with a shared-memory INPUT-TILE CACHE.
//
// jets_fused_fast_kernel streams every operand of every term through shared memory once per
// term, so a block-tridiagonal row d_r = w_r.*x_r + S1(x_{r+1}) + S2(x_{r-1}) pulls 4 tiles out
// of L2 for 3 algorithmic streams, and a 4x4 JopBlock of diagonals pulls each x block 4 times.
// The re-reads hit L2, but the L2->SM fabric (not DRAM) then bounds the kernel (measured: 0.78
// and 0.64 of HBM peak, profiles/r01_*).  Here one CTA walks a BUNDLE of consecutive output rows
// at the same tile position and keeps every input tile in a ring of shared-memory buffers (the
// "x ring") from its first to its last use: JetBlock_df!'s `_m = getblock(m, jblock)`
// (src/Jets.jl:1019) is fetched once per tile no matter how many rows read it, exactly as the
// adjoint's `_d` (:1044).  Operator state (diagonals, linearization points) streams through a
// second ring of slots, one slot per term group, as before.
//
//   producer warp   per (bundle, position) unit, per term group: waits for the state slot, waits
//                   for the x-ring buffers its first-use terms allocate (ring order; the planner
//                   guarantees the previous occupant's last use lies in an earlier group), posts
//                   ONE expect_tx on the slot's `full` mbarrier covering state + new input tiles
//                   and issues the cp.async.bulk copies.  Up to G lanes issue groups in parallel;
//                   the planner picks G so that no lane ever waits on a group of its own batch.
//   16 consumer warps  wait on the slot's `full` barrier only; evaluate the straight-line chains
//                   (fused_ops.cuh, same IEEE operations as the interpreter -> bit-identical),
//                   accumulate the row's terms in registers left to right, release the state slot
//                   and the x buffers whose last use this group was, store the row tile once.
// Tiles have a run-time length (<= the template's capacity) chosen by the planner so that the
// number of units is a multiple of the grid: no ragged last wave.
#include <type_traits>
#include "fused_ops.cuh"

namespace jets {
namespace {

constexpr int kPad = 16;
constexpr int kMaxRing = 16;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kHdr = 3 * 128 + kMaxRing * 80;          // sfull | sempty | xempty | slot metadata
constexpr int kHdrAligned = (kHdr + 127) & ~127;

// F_FLUSH: a marker slot without data -- this CTA has finished `nvalid` units that feed the cross-rank signals in
// bits 8-11 (also carried by the F_END sentinel)
enum : int { F_FIRST = 1, F_LAST = 2, F_END = 4, F_ACC = 8, F_BLK0 = 32, F_BLKEND = 64, F_FLUSH = 128 };

struct BMeta {                    // 80 bytes, 16B aligned
  char* out_tile;                 // absolute address of out[tile_start]
  const BGroupRec* rec;
  int32_t nvalid, flags, nterms, xrelease;   // xrelease: bit s -> arrive on xempty[s] after this group
  BTerm terms[kGroupTerms];       // xrel replaced by the ring slot
  int64_t pad2[2];
};
static_assert(sizeof(BMeta) == 80, "BMeta layout");

struct BundleParams {
  const BGroupRec* groups;
  const BundleRec* bundles;
  int32_t nbundles, NX, NS, sstreams, G, tile_elems;
  int64_t nunits;
  int64_t table_bytes;        // bytes of the plan tables starting at `groups` (prefetch bound)
  int32_t* sched;             // {next claim, finished CTAs, next early claim} or null (static round-robin)
  int64_t nclaims;            // dynamic claims in the launch (bundle-major, BundleRec::claim_begin / chunk)
  // the first `early_claims` claims (peer-memory stores of the distributed apply: NVLink-bound, not HBM-bound) are
  // taken from their own queue by the first `early_ctas` CTAs only, so that the other SMs stream the HBM-bound
  // units meanwhile instead of all queueing behind the link
  int32_t early_ctas;
  int64_t early_claims;
  unsigned long long* trace;  // JETS_B200_TRACE: 8 globaltimer stamps per CTA of the launch (null: off)
  int32_t no_first_static;    // A/B: every claim through the atomic counter
  int32_t pre_state;          // operator state may be fetched before griddepcontrol.wait (no kernel of the stream writes it)
  const char* in;
  char* out;
  int32_t hl, hr;
  int32_t axpby;              // store epilogue out = cA*acc + cO*out_old (coefficients below)
  ApplyCoef coef;
  // cross-rank gating of the distributed banded apply (dist.cu); gate.flags == nullptr: plain launch
  GateLaunch gate;
  int32_t sig_total[kGateFlags];
  int32_t sig_owned;
  int32_t* sig_done;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// Cross-rank flag words (distributed banded apply): written by a neighbouring GPU over NVLink into this
// rank's exchange arena, polled here; epochs only grow, so "flag >= value" is a signed difference.
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Timeline of one CTA (profiles/trace_launch.py): 0 entry, 1 barriers initialised, 2 producer past
// griddepcontrol.wait, 3 first tile landed (consumer), 4 first row tile stored, 5 producer issued its last group,
// 6 consumer saw the end sentinel (all stores issued), 7 state bytes put in flight before the wait.
// MODE 0 is the plain kernel; MODE 1 adds -- at compile time, so that the plain instantiation pays nothing for them
// (measured on one box: the same features as run-time branches cost the plain config-5 launch 4 %, config 1 10 %) --
// the cross-rank gating of the distributed apply (flag waits, flush markers, signals, alternative base pointers,
// exit wait) and the launch timeline.  MODE 2 is the plain kernel with the axpby store epilogue (out = cA*acc + cO*out,
// jets_apply_axpby): the producer prefetches the old output tile into L2 when it issues the row tile's last group and
// the consumers load it BEFORE they evaluate that group's terms -- read right before the store it is a dependent DRAM
// load per row tile (ncu on config 4's fused iteration: 64 us per launch for 268 MB, 0.64 of the copy peak).  In
// the gated modes the epilogue stays a run-time switch; MODE 0 does not know it.  MODE 3 = MODE 1 + the launch timeline
// (the per-slot "first tile / first store traced yet?" checks were 4.5 % of the consumers' stall samples in the gated launch).
#define JETS_AXPBY (MODE == 2 || ((MODE & 1) != 0 && P.axpby != 0))
#define JETS_TRACE(slot, val) do { if constexpr (MODE == 3) { if (P.trace) P.trace[(size_t)blockIdx.x * 8 + (slot)] = (val); } } while (0)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <typename T, int CW, int VPT, int MODE>
__global__ void __launch_bounds__(CW * 32 + 32, 1) jets_fused_bundle_kernel(const BundleParams P) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  constexpr int kConsumers = CW * 32;
  constexpr int kTileBytes = kConsumers * 16 * VPT;
  constexpr int kBufBytes = kTileBytes + 2 * kPad;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sm0 = smem_u32(smem);
  const uint32_t sfull0 = sm0, sempty0 = sm0 + 128, xempty0 = sm0 + 256;
  BMeta* meta = reinterpret_cast<BMeta*>(smem + 384);
  const int NX = P.NX, NS = P.NS;
  const uint32_t xring0 = sm0 + kHdrAligned;
  const uint32_t sring0 = xring0 + (uint32_t)NX * kBufBytes;
  const uint32_t slot_bytes = (uint32_t)P.sstreams * kBufBytes;
  const int tid = threadIdx.x;
  if (tid == 0) JETS_TRACE(0, gtime());
  if (tid < kMaxRing) {            // 16 threads initialise the three barrier arrays side by side
    mbar_init(sfull0 + 8 * tid, 1);
    mbar_init(sempty0 + 8 * tid, CW);
    mbar_init(xempty0 + 8 * tid, CW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Programmatic dependent launch: the next kernel of the stream may start its own prologue while
  // this grid drains; this grid's prologue (barrier init, plan-table fetch: immutable data) ran
  // while the previous grid drained.  Nothing the previous kernel may have written -- or may still
  // be reading -- is touched before griddepcontrol.wait.  The dependents are released AFTER this grid's own
  // wait (below): the next kernel then never overlaps the kernel BEFORE this one, so "what the previous launch
  // writes" (Context::pdl_out_lo/hi) is all the host has to know to allow operator-state loads before the wait --
  // and nothing is lost, its CTAs cannot become resident before this grid's CTAs exit anyway.
  if (tid >= kConsumers) {         // warm L2/L1 with the head of the plan tables (first records of the first bundle)
    const int64_t o = (int64_t)(tid - kConsumers) * 128;
    if (o < P.table_bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(P.groups) + o));
  }
  BundleRec B = P.bundles[0];
  __syncthreads();                 // barriers initialised; nothing a previous kernel may write has been touched yet
  if (tid == 0) JETS_TRACE(1, gtime());
  if (tid < kConsumers) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if constexpr ((MODE & 1) != 0) {
    if (P.gate.flags != nullptr && blockIdx.x == 0 && tid == 0) {
      // signals this launch owns but no unit feeds (e.g. "halo consumed" when no row reads that halo)
#pragma unroll
      for (int k = 0; k < kGateFlags; ++k)
        if (((P.sig_owned >> k) & 1) && P.sig_total[k] == 0 && P.gate.sig_addr[k] != nullptr) {
          __threadfence_system();
          st_release_sys(P.gate.sig_addr[k], P.gate.sig_val[k]);
        }
    }
  }

  if (tid >= kConsumers) {
    // =============================== producer warp ===============================
    const int lane = tid - kConsumers;
    const uint32_t lpad = P.hl ? kPad : 0, rpad = P.hr ? kPad : 0;
    const int64_t te = P.tile_elems;
    const int G = P.G;
    int slot = 0;
    uint32_t par = 0;
    uint32_t xbase = 0;            // x-ring allocations made by this CTA so far
    int b = 0;
    int waited = 0;                // gate flags this CTA has already seen raised (epochs only grow)
    int cur_sig = 0, sig_units = 0;   // signals the current bundle feeds / units of it this CTA has issued
    auto flush_marker = [&](int kind) {
      if (lane == 0) {
        mbar_wait(sempty0 + 8 * slot, par ^ 1);
        meta[slot].flags = kind | (cur_sig << BG_SIG_SHIFT);
        meta[slot].nvalid = sig_units;
        mbar_arrive(sfull0 + 8 * slot);
      }
      if (++slot == NS) { slot = 0; par ^= 1; }
      __syncwarp();
    };
    int64_t unit_end = B.unit_begin + B.npos;

    // One lane issues one term group of unit `q` into state slot `my`.
    // stage 0: the whole group.  stage 1 (before griddepcontrol.wait): only the operator-STATE streams, which no
    // kernel of this stream writes (the host checked) -- their bytes are announced with a plain expect_tx, the
    // barrier's one arrival stays pending.  stage 2: the rest of a group whose state is already in flight.
    auto issue = [&](const BGroupRec* rec, int64_t q, uint32_t xb, int my, uint32_t mypar, auto stage_tag) {
      constexpr int stage = decltype(stage_tag)::value;
      const int4 hd = __ldg(reinterpret_cast<const int4*>(&rec->nsstreams));  // nsstreams nterms xrel_mask flags
      const uint4 t01 = __ldg(reinterpret_cast<const uint4*>(&rec->terms[0]));
      const uint4 t23 = __ldg(reinterpret_cast<const uint4*>(&rec->terms[2]));
      const int64_t out_off = __ldg(&rec->out_off);
      const int64_t tile_start = (B.pos0 + (q - B.unit_begin)) * te;
      const int64_t rem = B.len - tile_start;
      const int nvalid = rem < te ? (int)rem : (int)te;
      const uint32_t bytes = lpad + (((uint32_t)nvalid * sizeof(T) + 15u) & ~15u) + rpad;
      const int64_t goff = tile_start * (int64_t)sizeof(T) - lpad;
      const int nss = hd.x, nterms = hd.y, xrel_mask = hd.z, gflags = hd.w;
      BTerm tt[kGroupTerms];
      *reinterpret_cast<uint4*>(&tt[0]) = t01;
      *reinterpret_cast<uint4*>(&tt[2]) = t23;
      // ring position of the unit's first allocation (one division per group, none per term)
      const uint32_t xb_use = xb / (uint32_t)NX;
      const uint32_t xb_mod = xb - xb_use * (uint32_t)NX;
      if constexpr (stage != 2) mbar_wait(sempty0 + 8 * my, mypar ^ 1);
      if constexpr (stage == 1) {
        if (nss > 0) {
          asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(sfull0 + 8 * my), "r"(bytes * (uint32_t)nss) : "memory");
          const uint32_t sb = sring0 + my * slot_bytes + (kPad - lpad);
          for (int k = 0; k < nss; ++k) {
            const char* src = reinterpret_cast<const char*>(__ldg(&rec->sptr[k]));
            bulk_g2s(sb + k * kBufBytes, src + goff, bytes, sfull0 + 8 * my);
          }
        }
        return;
      }
      (void)xrel_mask;
      uint32_t total = stage == 2 ? 0u : bytes * (uint32_t)nss;
      int xrelease = 0;
#pragma unroll
      for (int t = 0; t < kGroupTerms; ++t) {
        if (t < nterms) {
          // allocation index inside the unit = xdiv * NX + xmod (split by the planner)
          uint32_t xs = xb_mod + (uint32_t)(tt[t].xflags >> 4);
          uint32_t use = xb_use + ((tt[t].xflags >> 2) & 1u);
          if (xs >= (uint32_t)NX) { xs -= (uint32_t)NX; ++use; }
          if (tt[t].xflags & XF_LOAD) {
            mbar_wait(xempty0 + 8 * xs, (use & 1) ^ 1);
            total += bytes;
          }
          if (tt[t].xflags & XF_RELEASE) xrelease |= 1 << xs;
          tt[t].xrel = (uint16_t)xs;
        }
      }
      BMeta& M = meta[my];
      char* obase = P.out;
      if constexpr ((MODE & 1) != 0) {
        const int oalt = (gflags >> BG_OUT_ALT_SHIFT) & 3;
        obase = oalt == 0 ? P.out : oalt == 1 ? P.gate.out_alt[0] : oalt == 2 ? P.gate.out_alt[1] : P.gate.out_alt[2];
      }
      M.out_tile = obase + (out_off + tile_start) * (int64_t)sizeof(T);
      M.rec = rec;
      const int fl = (tile_start == 0 ? F_BLK0 : 0) | (rem <= te ? F_BLKEND : 0) |
                     ((gflags & BG_ROW_FIRST) ? F_FIRST : 0) | ((gflags & BG_ROW_LAST) ? F_LAST : 0) |
                     ((gflags & BG_ACC) ? F_ACC : 0);
      *reinterpret_cast<int4*>(&M.nvalid) = make_int4(nvalid, fl, nterms, xrelease);
      if constexpr (MODE == 2) {
        if ((fl & F_LAST) && nvalid * (int)sizeof(T) >= 16) bulk_prefetch_l2(M.out_tile, (uint32_t)(nvalid * (int)sizeof(T)) & ~15u);
      }
      *reinterpret_cast<uint4*>(&M.terms[0]) = *reinterpret_cast<uint4*>(&tt[0]);
      *reinterpret_cast<uint4*>(&M.terms[2]) = *reinterpret_cast<uint4*>(&tt[2]);
      if (total == 0) {
        mbar_arrive(sfull0 + 8 * my);
        return;
      }
      mbar_expect_tx(sfull0 + 8 * my, total);
#pragma unroll
      for (int t = 0; t < kGroupTerms; ++t) {
        if (t < nterms && (tt[t].xflags & XF_LOAD)) {
          const int64_t px = __ldg(&rec->xptr[t]);
          const char* ibase = P.in;
          if constexpr ((MODE & 1) != 0) {
            const int ialt = (xrel_mask >> (kXAltShift + 2 * t)) & 3;
            ibase = ialt == 0 ? P.in : ialt == 1 ? P.gate.in_alt[0] : ialt == 2 ? P.gate.in_alt[1] : P.gate.in_alt[2];
          }
          const char* src = ((xrel_mask >> t) & 1) ? ibase + px : reinterpret_cast<const char*>(px);
          bulk_g2s(xring0 + tt[t].xrel * kBufBytes + (kPad - lpad), src + goff, bytes, sfull0 + 8 * my);
        }
      }
      if constexpr (stage == 2) return;
      const uint32_t sb = sring0 + my * slot_bytes + (kPad - lpad);
      for (int k = 0; k < nss; ++k) {
        const char* src = reinterpret_cast<const char*>(__ldg(&rec->sptr[k]));
        bulk_g2s(sb + k * kBufBytes, src + goff, bytes, sfull0 + 8 * my);
      }
    };

    // Units are dealt either statically (unit q = blockIdx.x + k*gridDim.x) or, when the plan carries
    // a scheduler counter, dynamically in chunks of `chunk` consecutive units claimed with one atomic
    // (the next claim is issued while the current chunk is processed): CTAs that start late -- SMs
    // held by a concurrent NCCL kernel, by the previous grid's tail under programmatic launch --
    // simply take fewer chunks instead of stretching the tail.
    const bool dyn = P.sched != nullptr;
    const int64_t stride = dyn ? 1 : (int64_t)gridDim.x;
    // claim_issue() starts the atomic (lane 0 keeps the ticket in a register); claim_take() broadcasts
    // it when the next chunk is actually needed, so the atomic's latency hides behind a whole chunk.
    // Dynamic claims are enumerated per bundle (BundleRec::claim_begin / chunk: the planner sizes a claim for
    // roughly equal work, many light units or one heavy one); a claim never spans two bundles.
    int ticket = 0;
    const int64_t n_early = (dyn && P.early_ctas > 0) ? P.early_claims : 0;
    const int64_t ticket_base_v = (dyn && n_early == 0 && !P.no_first_static) ? (int64_t)gridDim.x : 0;
    int phase = (n_early > 0 && (int)blockIdx.x < P.early_ctas) ? 0 : 1;       // 0: draining the early queue
    auto claim_issue = [&]() { if (lane == 0) ticket = atomicAdd(P.sched + (phase == 0 ? 2 : 0), 1); };
    auto claim_take = [&]() -> int64_t {
      const int64_t t = (int64_t)__shfl_sync(0xffffffffu, ticket, 0);
      return phase == 0 ? t : n_early + t + ticket_base_v;
    };
    auto locate_claim = [&](int64_t c) {   // bundle of claim c (claims are bundle-major like units)
      if (c < B.claim_begin || c >= (int64_t)B.claim_begin + B.nclaims) {
        int lo = 0, hi = P.nbundles - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if ((int64_t)__ldg(&P.bundles[mid].claim_begin) <= c) lo = mid; else hi = mid - 1;
        }
        b = lo;
        B = P.bundles[b];
        unit_end = B.unit_begin + B.npos;
      }
    };
    int64_t q = blockIdx.x, q_end = P.nunits;
    int64_t c = 0;
    bool ticket_pending = false;           // the next ticket has been requested already
    auto open_claim = [&]() {              // unit range of claim c
      locate_claim(c);
      q = B.unit_begin + (c - B.claim_begin) * B.chunk;
      q_end = q + B.chunk < unit_end ? q + B.chunk : unit_end;
      ticket_pending = false;
    };
    auto locate_unit = [&]() {             // static mode: bundle of unit q (units are enumerated bundle-major)
      if (q >= unit_end || q < B.unit_begin) {
        int lo = 0, hi = P.nbundles - 1;
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (__ldg(&P.bundles[mid].unit_begin) <= q) lo = mid; else hi = mid - 1;
        }
        b = lo;
        B = P.bundles[b];
        unit_end = B.unit_begin + B.npos;
      }
    };
    // Without an early queue the FIRST claim of a CTA is static (claim blockIdx.x: the grid never exceeds the
    // number of claims), tickets count from gridDim.x: no atomic sits between the launch and the first load.
    const bool first_static = dyn && n_early == 0 && !P.no_first_static;
    bool have = true;
    int pre_groups = 0;                    // groups of the first unit whose state loads are already in flight
    if (first_static) { c = blockIdx.x; open_claim(); }
    else if (!dyn) locate_unit();
    if (P.pre_state && (first_static || !dyn) && q < q_end && B.gate == 0 && 2 * B.ngroups > G) {
      // the previous kernel of the stream is still draining: put the operator state of this CTA's first unit in
      // flight now (plan tables and operator state are immutable for it); everything else waits below
      pre_groups = B.ngroups < G ? B.ngroups : G;
      if (lane < pre_groups) issue(P.groups + B.group_begin + lane, q, 0u, lane < NS ? lane : lane - NS, 0u, std::integral_constant<int, 1>{});
      __syncwarp();
      if (lane == 0) JETS_TRACE(7, (unsigned long long)pre_groups);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0) JETS_TRACE(2, gtime());
    // The NEXT ticket is fetched while the last group batch of the current claim is being issued -- early enough to
    // hide the atomic, late enough not to reserve a unit this CTA will not start for a long time (traced on config 5:
    // with the ticket fetched when a claim was opened, every CTA held one unstarted 280 us unit in reserve when the
    // queue ran dry, and the CTAs finished 330 us apart).
    if (dyn && !first_static) {
      claim_issue();
      c = claim_take();
      if (phase == 0 && c >= n_early) { phase = 1; claim_issue(); c = claim_take(); }
      have = c < P.nclaims;
      if (have) open_claim();
    }
    while (have) {
      if (q >= q_end) {
        if (!dyn) break;
        c = claim_take();
        if (phase == 0 && c >= n_early) {      // early queue drained: join the main queue
          phase = 1;
          claim_issue();
          c = claim_take();
        }
        if (c >= P.nclaims) break;
        open_claim();
      }
      locate_unit();
      if constexpr ((MODE & 1) != 0) {
      if ((B.gate >> 4) != cur_sig) {
        // leaving a bundle whose units feed cross-rank signals: tell the consumers how many this CTA completed
        // (BEFORE any flag wait below -- a neighbour may be waiting for exactly this signal)
        if (cur_sig) flush_marker(F_FLUSH);
        cur_sig = B.gate >> 4;
        sig_units = 0;
      }
      if ((B.gate & 15) & ~waited) {
        // the unit reads (or overwrites) memory a neighbouring rank fills (or is still reading): wait for
        // its flag word.  Gated bundles are enumerated so that the waits are normally satisfied already.
        if (lane == 0) {
          int m = (B.gate & 15) & ~waited;
          while (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t* fp = P.gate.flags + k * kGateFlagStride;
            uint64_t t0 = 0;
            while ((int32_t)(ld_acquire_sys(fp) - P.gate.wait_val[k]) < 0) {
              __nanosleep(100);
              if (P.gate.timeout_ns) {
                uint64_t now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > P.gate.timeout_ns) { atomicAdd(P.gate.err, 1u); break; }
              }
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");   // the data arrived through the generic proxy; TMA reads it
        }
        waited |= B.gate & 15;
        __syncwarp();
      }
      }
      // Short bundles: several units of the bundle are issued side by side, one lane per group, as
      // long as the batch needs no x buffer that one of its own groups has to release first.
      int U = 1;
      if (2 * B.ngroups <= G) {
        U = G / B.ngroups;
        if (B.nx > 0 && U > NX / B.nx) U = NX / B.nx;
        const int64_t lim = unit_end < q_end ? unit_end : q_end;
        const int64_t mine = (lim - 1 - q) / stride + 1;   // units of this bundle this CTA still owns
        if (U > mine) U = (int)mine;
        if (U < 1) U = 1;
      }
      // Next ticket: light units (a claim of them is little work, and the producer runs through it in one or two
      // batches) request it up front so that the atomic hides behind the claim; heavy units request it in their last
      // group batch, see above.
      const bool last_of_claim = dyn && q + (int64_t)U * stride >= q_end;
      if (dyn && !ticket_pending && (last_of_claim || 2 * B.ngroups <= G) && B.ngroups <= G) { claim_issue(); ticket_pending = true; }
      if (U > 1) {
        const int n = U * B.ngroups;
        if (lane < n) {
          const int u = lane / B.ngroups, g = lane - u * B.ngroups;
          int my = slot + lane;
          uint32_t mypar = par;
          if (my >= NS) { my -= NS; mypar ^= 1; }
          issue(P.groups + B.group_begin + g, q + (int64_t)u * stride, xbase + (uint32_t)(u * B.nx), my, mypar, std::integral_constant<int, 0>{});
        }
        slot += n;
        if (slot >= NS) { slot -= NS; par ^= 1; }
        __syncwarp();
      } else {
        for (int g0 = 0; g0 < B.ngroups; g0 += G) {
          const int n = (B.ngroups - g0) < G ? (B.ngroups - g0) : G;
          if (last_of_claim && !ticket_pending && g0 + G >= B.ngroups) { claim_issue(); ticket_pending = true; }
          if (lane < n) {
            int my = slot + lane;
            uint32_t mypar = par;
            if (my >= NS) { my -= NS; mypar ^= 1; }
            if (g0 == 0 && lane < pre_groups) issue(P.groups + B.group_begin + g0 + lane, q, xbase, my, mypar, std::integral_constant<int, 2>{});
            else issue(P.groups + B.group_begin + g0 + lane, q, xbase, my, mypar, std::integral_constant<int, 0>{});
          }
          slot += n;
          if (slot >= NS) { slot -= NS; par ^= 1; }
          __syncwarp();
          pre_groups = 0;                 // only the very first batch of the CTA was issued ahead
        }
      }
      xbase += (uint32_t)(U * B.nx);
      q += (int64_t)U * stride;
      sig_units += U;
    }
    if (dyn && lane == 0) {   // the last CTA to run out of work re-arms the counters for the next launch
      __threadfence();
      if (atomicAdd(P.sched + 1, 1) == (int)gridDim.x - 1) {
        P.sched[0] = 0;
        P.sched[1] = 0;
        P.sched[2] = 0;
        __threadfence();
      }
    }
    if (lane == 0) JETS_TRACE(5, gtime());
    flush_marker(F_END);   // end-of-work sentinel (reports the last bundle's units as well)
    if constexpr ((MODE & 1) != 0)
    if (P.gate.exit_wait && lane == 0) {
      // the last CTA to run out of work keeps the grid alive until the neighbours have finished reading this
      // rank's input (their flag words): everything that follows on the stream may then overwrite it
      __threadfence();
      if (atomicAdd(P.sig_done + kGateFlags, 1) == (int)gridDim.x - 1) {
        P.sig_done[kGateFlags] = 0;
        int m = P.gate.exit_wait;
        while (m) {
          const int k = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t* fp = P.gate.flags + k * kGateFlagStride;
          uint64_t t0 = 0;
          while ((int32_t)(ld_acquire_sys(fp) - P.gate.exit_val[k]) < 0) {
            __nanosleep(200);
            if (P.gate.timeout_ns) {
              uint64_t now;
              asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
              if (t0 == 0) t0 = now;
              else if (now - t0 > P.gate.timeout_ns) { atomicAdd(P.gate.err, 1u); break; }
            }
          }
        }
      }
    }
  } else {
    // =============================== consumer warps ==============================
    T acc[VPT][V];
    int slot = 0;
    uint32_t par = 0;
    T cA = T(1), cO = T(0);
    if (JETS_AXPBY) {   // device scalars of the caller (read after griddepcontrol.wait: a previous kernel may produce them)
      double a = P.coef.a_ptr ? *P.coef.a_ptr : P.coef.a_const;
      if (P.coef.a_flags & JETS_COEF_INV) a = 1.0 / a;
      if (P.coef.a_flags & JETS_COEF_NEG) a = -a;
      double o = P.coef.o_ptr ? *P.coef.o_ptr : P.coef.o_const;
      if (P.coef.o_flags & JETS_COEF_INV) o = 1.0 / o;
      if (P.coef.o_flags & JETS_COEF_NEG) o = -o;
      cA = (T)a; cO = (T)o;
    }
    const unsigned char* xr_p = smem + kHdrAligned + kPad + tid * 16;                   // vector 0 of x buffer 0
    const unsigned char* sl_p = xr_p + (size_t)NX * kBufBytes;                          // vector 0, stream 0, slot 0
    bool tr_first = false, tr_store = false;
    if constexpr (MODE == 3) tr_first = tr_store = P.trace != nullptr && tid == 0;
    while (true) {
      mbar_wait(sfull0 + 8 * slot, par);
      if constexpr (MODE == 3) { if (tr_first) { JETS_TRACE(3, gtime()); tr_first = false; } }
      const BMeta& M = meta[slot];
      const int flags = M.flags;
      if constexpr ((MODE & 1) == 0) {
        if (flags & F_END) break;
      } else if (flags & (F_END | F_FLUSH)) {
        const int smask = (flags >> BG_SIG_SHIFT) & 15;
        const int cnt = M.nvalid;
        if (smask && cnt > 0) {
          // every consumer warp has issued its stores of those units (slots are consumed in order); one thread
          // orders them system-wide and counts; the CTA that completes a signal raises the neighbour's flag word
          // (data and flag may both live in peer memory: fence.sys orders them for the observer's acquire)
          asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
          if (tid == 0) {
            __threadfence_system();
            int m = smask;
            while (m) {
              const int k = __ffs(m) - 1;
              m &= m - 1;
              if (atomicAdd(P.sig_done + k, cnt) + cnt == P.sig_total[k]) {
                P.sig_done[k] = 0;               // re-armed for the next launch of this plan
                __threadfence_system();
                if (P.gate.sig_addr[k] != nullptr) st_release_sys(P.gate.sig_addr[k], P.gate.sig_val[k]);
              }
            }
          }
        }
        if (flags & F_END) {
          if (tid == 0) JETS_TRACE(6, gtime());
          break;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(sempty0 + 8 * slot);
        sl_p += slot_bytes;
        if (++slot == NS) { slot = 0; par ^= 1; sl_p -= (size_t)NS * slot_bytes; }
        continue;
      }
      const int nvalid = M.nvalid;
      const int nterms = M.nterms;
      const int xrelease = M.xrelease;
      T* out_tile = reinterpret_cast<T*>(M.out_tile);
      const CStage* stages = M.rec->stages;
      if (flags & F_FIRST) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int e0 = (i * kConsumers + tid) * V;
          if (flags & F_ACC) {
            if (e0 + V <= nvalid) {
              const Vec v = *reinterpret_cast<const Vec*>(out_tile + e0);
              const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
              for (int j = 0; j < V; ++j) acc[i][j] = vs[j];
            } else {
#pragma unroll
              for (int j = 0; j < V; ++j) acc[i][j] = (e0 + j < nvalid) ? out_tile[e0 + j] : T(0);
            }
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[i][j] = T(0);
          }
        }
      }
      auto run_terms = [&](auto edge_tag) {
        constexpr bool EDGE = decltype(edge_tag)::value;
        // per-vector edge information (constant across the terms of the group)
        bool first[VPT];
        int last[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int e0 = (i * kConsumers + tid) * V;
          first[i] = EDGE && (flags & F_BLK0) && e0 == 0;
          last[i] = (EDGE && (flags & F_BLKEND)) ? nvalid - 1 - e0 : (1 << 30);
        }
        for (int t = 0; t < nterms; ++t) {
          const BTerm gt = M.terms[t];
          const char* xin = reinterpret_cast<const char*>(xr_p) + (int)gt.xrel * kBufBytes;
          const char* sst = reinterpret_cast<const char*>(sl_p) + (int)gt.sstream0 * kBufBytes;
          FastIO2<T> io[VPT];
#pragma unroll
          for (int i = 0; i < VPT; ++i) {
            io[i].stride = kBufBytes;
            io[i].bin = xin + i * kConsumers * 16;
            io[i].bst = sst + i * kConsumers * 16;
          }
          T val[VPT][V];
          eval_fast_n<T, EDGE, VPT>(gt.pattern, io, stages + gt.stage0, first, last, val);   // one dispatch per term
          if (gt.sign >= 0) {
#pragma unroll
            for (int i = 0; i < VPT; ++i)
#pragma unroll
              for (int j = 0; j < V; ++j) acc[i][j] = acc[i][j] + val[i][j];
          } else {
#pragma unroll
            for (int i = 0; i < VPT; ++i)
#pragma unroll
              for (int j = 0; j < V; ++j) acc[i][j] = acc[i][j] - val[i][j];
          }
        }
      };
      Vec oldv[VPT];
      if constexpr (MODE == 2) {       // the old output tile, in flight while the terms are evaluated
        if (flags & F_LAST) {
#pragma unroll
          for (int i = 0; i < VPT; ++i) {
            const int e0 = (i * kConsumers + tid) * V;
            if (e0 + V <= nvalid) oldv[i] = *reinterpret_cast<const Vec*>(out_tile + e0);
          }
        }
      }
      // interior tiles (neither block end inside the tile) run the mask-free instantiation
      if (flags & (F_BLK0 | F_BLKEND)) run_terms(std::true_type{});
      else run_terms(std::false_type{});
      __syncwarp();
      if ((tid & 31) == 0) {
        mbar_arrive(sempty0 + 8 * slot);   // state slot may be refilled
        int m = xrelease;
        while (m) {                        // input tiles whose last use this group was
          const int s = __ffs(m) - 1;
          m &= m - 1;
          mbar_arrive(xempty0 + 8 * s);
        }
      }
      if (flags & F_LAST) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int e0 = (i * kConsumers + tid) * V;
          if (e0 + V <= nvalid) {
            Vec v;
            T* vs = reinterpret_cast<T*>(&v);
            if (JETS_AXPBY) {
              Vec old;
              if constexpr (MODE == 2) old = oldv[i];
              else old = *reinterpret_cast<const Vec*>(out_tile + e0);
              const T* os = reinterpret_cast<const T*>(&old);
#pragma unroll
              for (int j = 0; j < V; ++j) vs[j] = cA * acc[i][j] + cO * os[j];
            } else {
#pragma unroll
              for (int j = 0; j < V; ++j) vs[j] = acc[i][j];
            }
            *reinterpret_cast<Vec*>(out_tile + e0) = v;
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j)
              if (e0 + j < nvalid) out_tile[e0 + j] = JETS_AXPBY ? cA * acc[i][j] + cO * out_tile[e0 + j] : acc[i][j];
          }
        }
      }
      if constexpr (MODE == 3) { if (tr_store && (flags & F_LAST)) { JETS_TRACE(4, gtime()); tr_store = false; } }
      sl_p += slot_bytes;
      if (++slot == NS) { slot = 0; par ^= 1; sl_p -= (size_t)NS * slot_bytes; }
    }
  }
}

template <typename T, int CW, int VPT, int MODE>
void launch_variant(const DevFused& f, BundleParams& P, cudaStream_t s) {
  constexpr int kBufBytes = CW * 32 * 16 * VPT + 2 * kPad;
  const size_t smem = kHdrAligned + (size_t)f.NX * kBufBytes + (size_t)f.NS * f.sstreams * kBufBytes;
  JETS_CHECK(smem <= (size_t)kSmemLimit && f.NX >= 1 && f.NX <= kMaxRing && f.NS >= 1 && f.NS <= kMaxRing,
             JETS_ERR_INVALID, "internal: bundle kernel ring sizes NX=%d NS=%d do not fit", f.NX, f.NS);
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(jets_fused_bundle_kernel<T, CW, VPT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  int64_t grid = ctx().sm_count;
  if (ctx().grid_limit > 0 && ctx().grid_limit < grid) grid = ctx().grid_limit;
  const int64_t nclaims = P.sched ? P.nclaims : P.nunits;
  if (grid > nclaims) grid = nclaims > 0 ? nclaims : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(CW * 32 + 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx().no_pdl ? 0 : 1;
  CUDA_TRY(cudaLaunchKernelEx(&cfg, jets_fused_bundle_kernel<T, CW, VPT, MODE>, P));
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

template <typename T>
void launch_dtype(const DevFused& f, BundleParams& P, cudaStream_t s, bool extras) {
  if (P.axpby && !extras) {
    switch (f.variant) {
      case 2: launch_variant<T, 16, 2, 2>(f, P, s); return;
      case 0: launch_variant<T, 16, 1, 2>(f, P, s); return;
      default: JETS_FAIL(JETS_ERR_UNSUPPORTED, "the axpby store epilogue is built for tile shapes 0 and 2 only (got %d)", f.variant);
    }
  }
  if (extras && P.trace) {
    switch (f.variant) {
      case 2: launch_variant<T, 16, 2, 3>(f, P, s); return;
      case 0: launch_variant<T, 16, 1, 3>(f, P, s); return;
      default: JETS_FAIL(JETS_ERR_UNSUPPORTED, "the traced bundle kernel is built for tile shapes 0 and 2 only (got %d)", f.variant);
    }
  }
  if (extras) {
    // gated (distributed) launches: the tile shapes the planner picks for long rows and for short ones
    switch (f.variant) {
      case 2: launch_variant<T, 16, 2, 1>(f, P, s); return;
      case 0: launch_variant<T, 16, 1, 1>(f, P, s); return;
      default: JETS_FAIL(JETS_ERR_UNSUPPORTED, "the gated / traced bundle kernel is built for tile shapes 0 and 2 only (got %d)", f.variant);
    }
  }
  switch (f.variant) {
    case 1: launch_variant<T, 8, 2, 0>(f, P, s); break;
    case 2: launch_variant<T, 16, 2, 0>(f, P, s); break;
    case 3: launch_variant<T, 8, 4, 0>(f, P, s); break;
    case 4: launch_variant<T, 30, 1, 0>(f, P, s); break;
    case 5: launch_variant<T, 24, 1, 0>(f, P, s); break;
    default: launch_variant<T, 16, 1, 0>(f, P, s); break;
  }
}

}  // namespace

int bundle_buf_bytes(int variant) {
  static const int tile[6] = {8192, 8192, 16384, 16384, 15360, 12288};   // CW*32*16*VPT of launch_dtype's variants
  return tile[variant >= 0 && variant < 6 ? variant : 0] + 2 * kPad;
}
int bundle_smem_budget() { return kSmemLimit - kHdrAligned; }

void launch_fused_bundle(const DevFused& f, int dtype, const char* in, char* out, cudaStream_t s, const ApplyCoef* coef,
                         const GateLaunch* gate) {
  if (f.nbundles == 0 || f.nunits == 0) return;
  JETS_CHECK(!f.gated || (gate && gate->flags), JETS_ERR_INVALID, "internal: gated plan launched without its flag words");
  BundleParams P;
  P.axpby = coef ? 1 : 0;
  if (coef) P.coef = *coef;
  if (gate) P.gate = *gate;
  for (int k = 0; k < kGateFlags; ++k) P.sig_total[k] = f.sig_total[k];
  P.sig_owned = f.sig_owned;
  P.sig_done = f.sig_done;
  P.groups = f.bgroups; P.bundles = f.bundles;
  P.nbundles = f.nbundles; P.NX = f.NX; P.NS = f.NS; P.sstreams = f.sstreams; P.G = f.G;
  P.tile_elems = f.tile_elems; P.nunits = f.nunits;
  P.table_bytes = (int64_t)f.table_bytes;
  P.sched = f.sched;
  P.nclaims = f.nclaims;
  P.early_claims = f.early_claims;
  P.early_ctas = f.early_claims > 0 ? ctx().dist_early_ctas : 0;
  P.in = in; P.out = out; P.hl = f.hl; P.hr = f.hr;
  {
    // Operator state may be put in flight before griddepcontrol.wait unless the previous bundle launch -- the only
    // kernel that lets its dependents start early -- writes into it (e.g. a diagonal operator built on the vector
    // the previous apply produced).  Gated launches write through peer pointers: unknown, so never early after one.
    Context& c = ctx();
    const uintptr_t lo = reinterpret_cast<uintptr_t>(out) + (uintptr_t)f.out_lo, hi = reinterpret_cast<uintptr_t>(out) + (uintptr_t)f.out_hi;
    P.no_first_static = c.no_first_static;
    P.trace = nullptr;
    if (c.trace_buf) {          // ring of per-launch records: [launch index % kTraceLaunches][CTA][8]
      P.trace = c.trace_buf + (size_t)(c.trace_count % Context::kTraceLaunches) * Context::kTraceCtas * 8;
      ++c.trace_count;
    }
    P.pre_state = (!c.no_pdl && !c.no_pre_state && !gate && f.state_hi != 0 &&
                   !(f.state_lo < c.pdl_out_hi && c.pdl_out_lo < f.state_hi)) ? 1 : 0;
    if (gate) { c.pdl_out_lo = 0; c.pdl_out_hi = ~(uintptr_t)0; }
    else { c.pdl_out_lo = lo; c.pdl_out_hi = hi; }
  }
  // the plain kernel unless the launch is gated or traced (a traced launch with another tile shape stays untraced)
  const bool extras = gate != nullptr || (P.trace != nullptr && (f.variant == 0 || f.variant == 2));
  if (!extras) P.trace = nullptr;
  if (dtype == JETS_F32) launch_dtype<float>(f, P, s, extras);
  else launch_dtype<double>(f, P, s, extras);
}

}  // namespace jets
