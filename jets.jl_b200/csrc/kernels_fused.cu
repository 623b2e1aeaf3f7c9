// The fused block-apply engines: ONE launch walks a device-side table of output block rows;
// every row is a signed sum of "terms" (an elementwise/stencil chain applied to one input
// block), so JopBlock row sums, JopSum terms and JopComposite chains are all evaluated in
// registers and each output element is stored exactly once
// (replaces JetBlock_df!/df'! src/Jets.jl:1010-1057, JetSum_* :630-655, JetComposite_* :524-540
// and the leaf broadcasts of test/runtests.jl:3-4,20-21).
//
// TMA engine (default): a warp-specialised persistent kernel, one CTA per SM.  A producer warp
// streams operand tiles HBM -> shared memory with cp.async.bulk (1-D TMA) through a ring of
// mbarrier-guarded slots; one slot carries every operand stream of a GROUP of terms of one output
// tile (up to 4 streams), so a block-tridiagonal row (3 terms, 4 streams) costs one barrier round
// trip per tile.  16 consumer warps interpret the chains on 128-bit vectors read from shared
// memory (stencil halo elements come from the same staged tile) and write the output with
// coalesced 128-bit stores.
// Tiles are visited in (super-chunk, row, position) order and dealt round-robin to the CTAs: at
// any instant all SMs stream one contiguous window of a block (DRAM-page friendly, like a plain
// streaming kernel), while rows that share an input block follow each other closely enough for the
// shared tile to be an L2 hit.
// LDG engine (fallback for unaligned / caller-owned memory, and the A/B comparison): same
// interpreter, operands read with guarded global loads.
#include "fused_ops.cuh"

namespace jets {
namespace {

constexpr int kTileBytes = 8192;            // per stream per slot
constexpr int kPad = 16;                    // halo padding on either side of a staged tile
constexpr int kBufBytes = kTileBytes + 2 * kPad;
constexpr int kConsumerWarps = 16;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kThreads = kConsumers + 32;   // + producer warp
constexpr int kMaxSlots = 16;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kLdgThreads = 256;
constexpr int kLdgVPT = kTileBytes / 16 / kLdgThreads;

enum : int { F_FIRST = 1, F_LAST = 2, F_END = 4, F_ACC = 8, F_NOTERM = 16 };

struct SlotMeta {                 // written by the producer, read by consumers (smem), 48 bytes
  int64_t tile_start;             // block-local element index of the tile
  int64_t len;                    // block length
  char* out_tile;                 // absolute address of out[tile_start]
  const GroupRec* rec;            // static description of the term group (global, read-only)
  int32_t nvalid;
  int32_t flags;
  int32_t nterms;
  int32_t pad;
  GTerm terms[kGroupTerms];       // copied from rec so the consumers never wait on global memory
};

constexpr int kHdrBytes = 256 + kMaxSlots * (int)sizeof(SlotMeta);
constexpr int kHdrAligned = (kHdrBytes + 127) & ~127;

struct FusedParams {
  const FStage* stages;
  const FTerm* terms;
  const GroupRec* groups;
  const FRow* rows;
  const FSeg* segs;
  const int32_t* order;
  int32_t nsegs;
  int32_t nslots;
  int32_t slot_streams;
  int32_t S;          // tiles per (super-chunk, row) item
  int64_t nitems;
  const char* in;
  char* out;
};

// ------------------------------------------------------------------ PTX helpers ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA: global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ------------------------------------------------------------------ schedule -------------
// item -> (row, super-chunk): items are ordered super-chunk-major so that all rows touching the
// same positions of an input block run back to back (the shared input tile is an L2 hit).
__device__ __forceinline__ void decode_item(const FusedParams& P, int64_t item, int32_t& row,
                                            int64_t& chunk) {
  int lo = 0, hi = P.nsegs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (P.segs[mid].tile_begin <= item) lo = mid; else hi = mid - 1;
  }
  const FSeg sg = P.segs[lo];
  const int64_t t = item - sg.tile_begin;
  chunk = sg.pos_begin + t / sg.nactive;
  row = P.order[(int)(t % sg.nactive)];
}

// ------------------------------------------------------------------ loaders --------------
template <typename T, int HL, int HR, int NV>
struct SmemLoader {
  using Vec = typename VecOf<T>::type;
  static constexpr int V = VecOf<T>::V;
  static constexpr int W = HL + V + HR;
  const char* base;  // slot base + stream0 of the current term
  int e0[NV];        // element offset of each owned vector inside the tile
  __device__ __forceinline__ void operator()(int k, int i, T (&out)[W]) const {
    const char* b = base + k * kBufBytes + kPad + e0[i] * (int)sizeof(T);
    const Vec v = *reinterpret_cast<const Vec*>(b);
    const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) out[HL + j] = vs[j];
#pragma unroll
    for (int j = 0; j < HL; ++j) out[j] = *reinterpret_cast<const T*>(b - (HL - j) * (int)sizeof(T));
#pragma unroll
    for (int j = 0; j < HR; ++j) out[HL + V + j] = *reinterpret_cast<const T*>(b + (V + j) * (int)sizeof(T));
  }
};

template <typename T> __device__ __forceinline__ T ldg1(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ Cx<float> ldg1(const Cx<float>* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return Cx<float>(v.x, v.y);
}
template <> __device__ __forceinline__ Cx<double> ldg1(const Cx<double>* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return Cx<double>(v.x, v.y);
}

template <typename T, int HL, int HR, int NV>
struct GlobalLoader {
  using Vec = typename VecOf<T>::type;
  static constexpr int V = VecOf<T>::V;
  static constexpr int W = HL + V + HR;
  const char* ptr[kMaxStreams];  // block start of each stream
  int64_t p0[NV];
  int64_t len;
  __device__ __forceinline__ void operator()(int k, int i, T (&out)[W]) const {
    const T* base = reinterpret_cast<const T*>(ptr[k]);
    const int64_t p = p0[i];
    if (p + V <= len && ((reinterpret_cast<uintptr_t>(base + p) & 15) == 0)) {
      const Vec v = __ldg(reinterpret_cast<const Vec*>(base + p));
      const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int j = 0; j < V; ++j) out[HL + j] = vs[j];
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) out[HL + j] = (p + j < len) ? ldg1(base + p + j) : T(0);
    }
#pragma unroll
    for (int j = 0; j < HL; ++j) {
      const int64_t q = p - HL + j;
      out[j] = (q >= 0 && q < len) ? ldg1(base + q) : T(0);
    }
#pragma unroll
    for (int j = 0; j < HR; ++j) {
      const int64_t q = p + V + j;
      out[HL + V + j] = (q < len) ? ldg1(base + q) : T(0);
    }
  }
};

template <typename T>
__device__ __forceinline__ void load_out(const char* out_tile, int e0, int nvalid,
                                         T (&acc)[VecOf<T>::V]) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  const T* o = reinterpret_cast<const T*>(out_tile) + e0;
  if (e0 + V <= nvalid && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
    const Vec v = *reinterpret_cast<const Vec*>(o);
    const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = vs[j];
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = (e0 + j < nvalid) ? o[j] : T(0);
  }
}
template <typename T>
__device__ __forceinline__ void store_out(char* out_tile, int e0, int nvalid,
                                          const T (&acc)[VecOf<T>::V]) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  T* o = reinterpret_cast<T*>(out_tile) + e0;
  if (e0 + V <= nvalid && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
    Vec v;
    T* vs = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) vs[j] = acc[j];
    *reinterpret_cast<Vec*>(o) = v;
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (e0 + j < nvalid) o[j] = acc[j];
  }
}

// ------------------------------------------------------------------ TMA engine -----------
template <typename T, int HL, int HR, bool HEAVY>
__global__ void __launch_bounds__(kThreads, 1) jets_fused_tma_kernel(const FusedParams P) {
  constexpr int V = VecOf<T>::V;
  constexpr int W = HL + V + HR;
  constexpr int kTileElems = kTileBytes / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kMaxSlots]
  uint64_t* empty = full + kMaxSlots;                          // [kMaxSlots]
  SlotMeta* meta = reinterpret_cast<SlotMeta*>(smem + 256);    // [kMaxSlots]
  unsigned char* slots = smem + kHdrAligned;
  const int nslots = P.nslots;
  const int slot_bytes = P.slot_streams * kBufBytes;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kConsumers) {
    // =============================== producer warp ===============================
    // Tiles q = blockIdx.x, +gridDim.x, ... ; q = item * S + tin.  Lane g of the warp issues the
    // g-th term group of the current tile (its own slot, its own barrier, its own bulk copies).
    const int lane = tid - kConsumers;
    const int64_t Q = P.nitems * P.S;
    int slot = 0;            // next slot to fill (ring position) and its use parity
    uint32_t par = 0;
    int64_t item = (int64_t)blockIdx.x / P.S;
    int tin = (int)((int64_t)blockIdx.x - item * P.S);
    int64_t cur_item = -1;
    int64_t chunk = 0;
    int32_t row_id = 0;
    FRow row;
    int ngroups = 0;
    const int G = nslots < 32 ? nslots : 32;
    for (int64_t q = blockIdx.x; q < Q; q += gridDim.x) {
      if (item != cur_item) {
        cur_item = item;
        decode_item(P, item, row_id, chunk);
        row = P.rows[row_id];
        ngroups = row.group_end - row.group_begin;
      }
      const int64_t pos = chunk * P.S + tin;
      tin += (int)gridDim.x;
      while (tin >= P.S) { tin -= P.S; ++item; }
      if (pos >= row.ntiles) continue;  // phantom tile of a ragged last super-chunk
      const int64_t tile_start = pos * kTileElems;
      const int64_t rem = row.len - tile_start;
      const int nvalid = rem < kTileElems ? (int)rem : kTileElems;
      const uint32_t bytes = (HL ? kPad : 0) + (((uint32_t)nvalid * sizeof(T) + 15u) & ~15u) + (HR ? kPad : 0);
      char* out_tile = P.out + (row.out_off + tile_start) * (int64_t)sizeof(T);
      const int nissue = ngroups > 0 ? ngroups : 1;
      for (int g0 = 0; g0 < nissue; g0 += G) {
        const int g = g0 + lane;
        const int n = (nissue - g0) < G ? (nissue - g0) : G;
        if (lane < n) {
          int my = slot + lane;
          uint32_t mypar = par;
          if (my >= nslots) { my -= nslots; mypar ^= 1; }
          const GroupRec* rec = P.groups + row.group_begin + g;
          int nstreams = 0, rel = 0, nterms = 0;
          uint4 t01 = make_uint4(0, 0, 0, 0), t23 = make_uint4(0, 0, 0, 0);
          if (ngroups > 0) {
            const int4 hd = __ldg(reinterpret_cast<const int4*>(&rec->nstreams));
            nstreams = hd.x; nterms = hd.y; rel = hd.z;
            t01 = __ldg(reinterpret_cast<const uint4*>(&rec->terms[0]));
            t23 = __ldg(reinterpret_cast<const uint4*>(&rec->terms[2]));
          }
          mbar_wait(&empty[my], mypar ^ 1);
          SlotMeta& M = meta[my];
          M.tile_start = tile_start;
          M.len = row.len;
          M.out_tile = out_tile;
          M.rec = rec;
          M.nvalid = nvalid;
          M.nterms = nterms;
          *reinterpret_cast<uint4*>(&M.terms[0]) = t01;
          *reinterpret_cast<uint4*>(&M.terms[2]) = t23;
          const int fl = (g == 0 ? F_FIRST : 0) | (g == nissue - 1 ? F_LAST : 0) | (row.init == 1 ? F_ACC : 0);
          if (ngroups == 0) {
            M.flags = fl | F_NOTERM;
            mbar_arrive(&full[my]);
          } else {
            M.flags = fl;
            mbar_expect_tx(&full[my], bytes * nstreams);
            unsigned char* sb = slots + (size_t)my * slot_bytes + (HL ? 0 : kPad);
            const int64_t goff = tile_start * (int64_t)sizeof(T) - (HL ? kPad : 0);
            for (int k = 0; k < nstreams; ++k) {
              const int64_t pk = __ldg(&rec->ptr[k]);
              const char* src = ((rel >> k) & 1) ? P.in + pk : reinterpret_cast<const char*>(pk);
              bulk_g2s(sb + k * kBufBytes, src + goff, bytes, &full[my]);
            }
          }
        }
        slot += n;
        if (slot >= nslots) { slot -= nslots; par ^= 1; }
        __syncwarp();
      }
    }
    if (lane == 0) {  // end-of-work sentinel
      mbar_wait(&empty[slot], par ^ 1);
      meta[slot].flags = F_END;
      mbar_arrive(&full[slot]);
    }
  } else {
    // =============================== consumer warps ==============================
    T acc[V];
    SmemLoader<T, HL, HR, 1> ld;
    ld.e0[0] = tid * V;
    int slot = 0;
    uint32_t par = 0;
    while (true) {
      mbar_wait(&full[slot], par);
      const SlotMeta& M = meta[slot];
      const int flags = M.flags;
      if (flags & F_END) break;
      const int nvalid = M.nvalid;
      char* out_tile = M.out_tile;
      if (flags & F_FIRST) {
        if (flags & F_ACC) load_out<T>(out_tile, ld.e0[0], nvalid, acc);
        else {
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] = T(0);
        }
      }
      if (!(flags & F_NOTERM)) {
        const GroupRec* __restrict__ rec = M.rec;
        const int nterms = M.nterms;
        const int64_t len = M.len;
        int64_t p0[1];
        p0[0] = M.tile_start + ld.e0[0];
        const bool first = (p0[0] == 0);
        const int64_t rl = len - 1 - p0[0];
        const int last = rl > V ? V : (rl < -1 ? -1 : (int)rl);
        const char* sbase = reinterpret_cast<const char*>(slots + (size_t)slot * slot_bytes);
        for (int t = 0; t < nterms; ++t) {
          const GTerm gt = M.terms[t];
          T val[V];
          FastIO<T> io;
          io.b0 = sbase + gt.stream0 * kBufBytes + kPad + ld.e0[0] * (int)sizeof(T);
          io.stride = kBufBytes;
          bool done = false;
          if constexpr (!IsCx<T>::value) done = eval_fast<T>(gt.pattern, io, rec->stages + gt.stage0, first, last, val);
          if (!done) {     // generic chains, and every chain of a complex space (conj-aware stages live in the interpreter)
            ld.base = sbase + gt.stream0 * kBufBytes;
            T wv[1][W];
            eval_term<T, HL, HR, 1, HEAVY>(rec->stages + gt.stage0, gt.nstages, ld, p0, len, wv);
#pragma unroll
            for (int j = 0; j < V; ++j) val[j] = wv[0][HL + j];
          }
          if (gt.sign >= 0) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = acc[j] + val[j];
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = acc[j] - val[j];
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&empty[slot]);  // slot may be refilled
      if (flags & F_LAST) store_out<T>(out_tile, ld.e0[0], nvalid, acc);
      if (++slot == nslots) { slot = 0; par ^= 1; }
    }
  }
}

// ------------------------------------------------------------------ LDG engine -----------
template <typename T, int HL, int HR, bool HEAVY>
__global__ void __launch_bounds__(kLdgThreads) jets_fused_ldg_kernel(const FusedParams P) {
  constexpr int V = VecOf<T>::V;
  constexpr int W = HL + V + HR;
  constexpr int kTileElems = kTileBytes / (int)sizeof(T);
  const int tid = threadIdx.x;
  const int64_t Q = P.nitems * P.S;
  int64_t cur_item = -1;
  int64_t chunk = 0;
  int32_t row_id = 0;
  FRow row;
  for (int64_t q = blockIdx.x; q < Q; q += gridDim.x) {
    const int64_t item = q / P.S;
    const int tin = (int)(q - item * P.S);
    if (item != cur_item) {
      cur_item = item;
      decode_item(P, item, row_id, chunk);
      row = P.rows[row_id];
    }
    const int64_t pos = chunk * P.S + tin;
    if (pos >= row.ntiles) continue;
    const int nterms = row.term_end - row.term_begin;
    const int64_t tile_start = pos * kTileElems;
    const int64_t rem = row.len - tile_start;
    const int nvalid = rem < kTileElems ? (int)rem : kTileElems;
    char* out_tile = P.out + (row.out_off + tile_start) * (int64_t)sizeof(T);
    T acc[kLdgVPT][V];
    GlobalLoader<T, HL, HR, kLdgVPT> ld;
    ld.len = row.len;
    int e0[kLdgVPT];
#pragma unroll
    for (int i = 0; i < kLdgVPT; ++i) {
      e0[i] = (i * kLdgThreads + tid) * V;
      ld.p0[i] = tile_start + e0[i];
      if (row.init == 1) load_out<T>(out_tile, e0[i], nvalid, acc[i]);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) acc[i][j] = T(0);
      }
    }
    for (int t = 0; t < nterms; ++t) {
      const FTerm tm = P.terms[row.term_begin + t];
      CStage cs[kMaxStages];
      ld.ptr[0] = tm.in_abs ? reinterpret_cast<const char*>(tm.in_abs)
                            : P.in + tm.in_off * (int64_t)sizeof(T);
      int k = 1;
      const int ns = tm.stage_end - tm.stage_begin;
#pragma unroll
      for (int s = 0; s < kMaxStages; ++s) {
        if (s < ns) {
          const FStage fs = P.stages[tm.stage_begin + s];
          cs[s].op = (uint8_t)fs.op; cs[s].fn = (uint8_t)fs.fn; cs[s].c0 = fs.c0;
          cs[s].has_stream = fs.ptr != nullptr;
          if (fs.ptr) {
#pragma unroll
            for (int kk = 1; kk < kMaxStreams; ++kk)
              if (kk == k) ld.ptr[kk] = reinterpret_cast<const char*>(fs.ptr);
            ++k;
          }
        }
      }
      T val[kLdgVPT][W];
      eval_term<T, HL, HR, kLdgVPT, HEAVY>(cs, ns, ld, ld.p0, row.len, val);
      if (tm.sign >= 0) {
#pragma unroll
        for (int i = 0; i < kLdgVPT; ++i)
#pragma unroll
          for (int j = 0; j < V; ++j) acc[i][j] = acc[i][j] + val[i][HL + j];
      } else {
#pragma unroll
        for (int i = 0; i < kLdgVPT; ++i)
#pragma unroll
          for (int j = 0; j < V; ++j) acc[i][j] = acc[i][j] - val[i][HL + j];
      }
    }
#pragma unroll
    for (int i = 0; i < kLdgVPT; ++i) store_out<T>(out_tile, e0[i], nvalid, acc[i]);
  }
}

template <typename T, int HL, int HR, bool HEAVY>
void launch_one(const DevFused& f, const FusedParams& P, cudaStream_t s) {
  const int sms = ctx().sm_count;
  const int64_t Q = P.nitems * P.S;
  if (f.use_tma) {
    const size_t smem = kHdrAligned + (size_t)P.nslots * P.slot_streams * kBufBytes;
    static bool attr_set = false;
    if (!attr_set) {
      CUDA_TRY(cudaFuncSetAttribute(jets_fused_tma_kernel<T, HL, HR, HEAVY>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
      attr_set = true;
    }
    int64_t grid = sms;
    if (grid > Q) grid = Q > 0 ? Q : 1;
    jets_fused_tma_kernel<T, HL, HR, HEAVY><<<(unsigned)grid, kThreads, smem, s>>>(P);
  } else {
    int64_t grid = (int64_t)sms * 8;
    if (grid > Q) grid = Q > 0 ? Q : 1;
    jets_fused_ldg_kernel<T, HL, HR, HEAVY><<<(unsigned)grid, kLdgThreads, 0, s>>>(P);
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

// Complex spaces (interleaved re/im): the interpreter kernels, TMA-staged like the real ones when every stream is
// aligned and guarded (a ComplexF64 halo element is exactly the 16-byte pad of a staged tile), guarded loads otherwise.
// The straight-line pattern kernels (fast / bundle) are not instantiated for them; x^2 is the only pointwise function.
template <typename T>
void launch_halo_cplx(const DevFused& f, const FusedParams& P, cudaStream_t s) {
  JETS_CHECK(!f.heavy, JETS_ERR_UNSUPPORTED, "internal: complex plans know x^2 only");
  if (f.hl == 0 && f.hr == 0) launch_one<T, 0, 0, false>(f, P, s);
  else if (f.hl == 0 && f.hr == 1) launch_one<T, 0, 1, false>(f, P, s);
  else if (f.hl == 1 && f.hr == 0) launch_one<T, 1, 0, false>(f, P, s);
  else if (f.hl == 1 && f.hr == 1) launch_one<T, 1, 1, false>(f, P, s);
  else JETS_FAIL(JETS_ERR_UNSUPPORTED, "fused halo (%d,%d) not instantiated", f.hl, f.hr);
}

template <typename T, bool HEAVY>
void launch_halo(const DevFused& f, const FusedParams& P, cudaStream_t s) {
  if (f.hl == 0 && f.hr == 0) launch_one<T, 0, 0, HEAVY>(f, P, s);
  else if (f.hl == 0 && f.hr == 1) launch_one<T, 0, 1, HEAVY>(f, P, s);
  else if (f.hl == 1 && f.hr == 0) launch_one<T, 1, 0, HEAVY>(f, P, s);
  else if (f.hl == 1 && f.hr == 1) launch_one<T, 1, 1, HEAVY>(f, P, s);
  else JETS_FAIL(JETS_ERR_UNSUPPORTED, "fused halo (%d,%d) not instantiated", f.hl, f.hr);
}

}  // namespace

int fused_tile_elems(int dtype) { return kTileBytes / (int)dsize(dtype); }

int fused_nslots(int slot_streams) {
  int n = (kSmemLimit - kHdrAligned) / (slot_streams * kBufBytes);
  if (n > kMaxSlots) n = kMaxSlots;
  return n;
}

void launch_fused(const DevFused& f, int dtype, const char* in, char* out, cudaStream_t s) {
  if (f.nrows == 0 || f.ntiles == 0) return;
  FusedParams P;
  P.stages = f.stages; P.terms = f.terms; P.groups = f.groups; P.rows = f.rows; P.segs = f.segs;
  P.order = f.order;
  P.nsegs = f.nsegs; P.nslots = fused_nslots(f.slot_streams); P.slot_streams = f.slot_streams;
  P.S = f.S; P.nitems = f.nitems; P.in = in; P.out = out;
  if (dtype == JETS_C64) launch_halo_cplx<Cx<float>>(f, P, s);
  else if (dtype == JETS_C128) launch_halo_cplx<Cx<double>>(f, P, s);
  else if (dtype == JETS_F32) {
    if (f.heavy) launch_halo<float, true>(f, P, s); else launch_halo<float, false>(f, P, s);
  } else {
    if (f.heavy) launch_halo<double, true>(f, P, s); else launch_halo<double, false>(f, P, s);
  }
}

}  // namespace jets
