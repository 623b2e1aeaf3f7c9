// Fast TMA engine: the warp-specialised streaming kernel for plans whose every chain has a
// straight-line fast path (fused_ops.cuh eval_fast: JopBlock of diagonal/stencil blocks, the
// D ∘ S ∘ J composite, B - c*S sums, ...).  Same slot/barrier protocol and schedule as
// jets_fused_tma_kernel (kernels_fused.cu), minus the stage interpreter, with the per-slot scalar
// work trimmed: 32-bit shared addresses, block-edge masks only on the vectors that touch an
// edge, whole-slot metadata fetched with 128-bit shared loads.
//
// Template: CW consumer warps, VPT 128-bit vectors per consumer thread per tile
// (tile = CW*32*16*VPT bytes per operand stream).
#include "fused_ops.cuh"

namespace jets {
namespace {

constexpr int kPad = 16;
constexpr int kMaxSlots = 16;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kHdr = 256 + kMaxSlots * 80;            // barriers + slot metadata
constexpr int kHdrAligned = (kHdr + 127) & ~127;

enum : int { F_FIRST = 1, F_LAST = 2, F_END = 4, F_ACC = 8, F_NOTERM = 16, F_BLK0 = 32, F_BLKEND = 64 };

struct FastMeta {                 // 80 bytes, 16B aligned: 5 x LDS.128 on the consumer side
  char* out_tile;                 // absolute address of out[tile_start]
  const GroupRec* rec;
  int32_t nvalid, flags, nterms, pad;
  GTerm terms[kGroupTerms];
  int64_t pad2[2];
};
static_assert(sizeof(FastMeta) == 80, "FastMeta layout");

struct FastParams {
  const GroupRec* groups;
  const FRow* rows;
  const FSeg* segs;
  const int32_t* order;
  int32_t nsegs, nslots, slot_streams, S;
  int64_t nitems;
  const char* in;
  char* out;
  int32_t hl, hr;                 // halo bytes to stage on each side are 16*hl / 16*hr
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                   "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void decode_item(const FastParams& P, int64_t item, int32_t& row, int64_t& chunk) {
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

template <typename T, int CW, int VPT>
__global__ void __launch_bounds__(CW * 32 + 32, 1) jets_fused_fast_kernel(const FastParams P) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  constexpr int kConsumers = CW * 32;
  constexpr int kTileBytes = kConsumers * 16 * VPT;
  constexpr int kBufBytes = kTileBytes + 2 * kPad;
  constexpr int kTileElems = kTileBytes / (int)sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sm0 = smem_u32(smem);
  const uint32_t full0 = sm0, empty0 = sm0 + 128, slots0 = sm0 + kHdrAligned;
  FastMeta* meta = reinterpret_cast<FastMeta*>(smem + 256);
  const int nslots = P.nslots;
  const uint32_t slot_bytes = (uint32_t)P.slot_streams * kBufBytes;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kConsumers) {
    // =============================== producer warp ===============================
    const int lane = tid - kConsumers;
    const int64_t Q = P.nitems * P.S;
    const uint32_t lpad = P.hl ? kPad : 0, rpad = P.hr ? kPad : 0;
    int slot = 0;
    uint32_t par = 0;
    int64_t item = (int64_t)blockIdx.x / P.S;
    int tin = (int)((int64_t)blockIdx.x - item * P.S);
    int64_t cur_item = -1, chunk = 0;
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
      const uint32_t bytes = lpad + (((uint32_t)nvalid * sizeof(T) + 15u) & ~15u) + rpad;
      char* out_tile = P.out + (row.out_off + tile_start) * (int64_t)sizeof(T);
      const int tflags = (row.init == 1 ? F_ACC : 0) | (tile_start == 0 ? F_BLK0 : 0) | (rem <= kTileElems ? F_BLKEND : 0);
      const int nissue = ngroups > 0 ? ngroups : 1;
      for (int g0 = 0; g0 < nissue; g0 += G) {
        const int g = g0 + lane;
        const int n = (nissue - g0) < G ? (nissue - g0) : G;
        if (lane < n) {
          int my = slot + lane;
          uint32_t mypar = par;
          if (my >= nslots) { my -= nslots; mypar ^= 1; }
          const GroupRec* rec = P.groups + row.group_begin + g;
          int4 hd = make_int4(0, 0, 0, 0);
          uint4 t01 = make_uint4(0, 0, 0, 0), t23 = make_uint4(0, 0, 0, 0);
          if (ngroups > 0) {
            hd = __ldg(reinterpret_cast<const int4*>(&rec->nstreams));
            t01 = __ldg(reinterpret_cast<const uint4*>(&rec->terms[0]));
            t23 = __ldg(reinterpret_cast<const uint4*>(&rec->terms[2]));
          }
          const int nstreams = hd.x, rel = hd.z;
          mbar_wait(empty0 + 8 * my, mypar ^ 1);
          FastMeta& M = meta[my];
          M.out_tile = out_tile;
          M.rec = rec;
          const int fl = tflags | (g == 0 ? F_FIRST : 0) | (g == nissue - 1 ? F_LAST : 0) | (ngroups == 0 ? F_NOTERM : 0);
          *reinterpret_cast<int4*>(&M.nvalid) = make_int4(nvalid, fl, hd.y, 0);
          *reinterpret_cast<uint4*>(&M.terms[0]) = t01;
          *reinterpret_cast<uint4*>(&M.terms[2]) = t23;
          if (ngroups == 0) {
            mbar_arrive(full0 + 8 * my);
          } else {
            mbar_expect_tx(full0 + 8 * my, bytes * nstreams);
            const uint32_t sb = slots0 + my * slot_bytes + (kPad - lpad);
            const int64_t goff = tile_start * (int64_t)sizeof(T) - lpad;
            for (int k = 0; k < nstreams; ++k) {
              const int64_t pk = __ldg(&rec->ptr[k]);
              const char* src = ((rel >> k) & 1) ? P.in + pk : reinterpret_cast<const char*>(pk);
              bulk_g2s(sb + k * kBufBytes, src + goff, bytes, full0 + 8 * my);
            }
          }
        }
        slot += n;
        if (slot >= nslots) { slot -= nslots; par ^= 1; }
        __syncwarp();
      }
    }
    if (lane == 0) {  // end-of-work sentinel
      mbar_wait(empty0 + 8 * slot, par ^ 1);
      meta[slot].flags = F_END;
      mbar_arrive(full0 + 8 * slot);
    }
  } else {
    // =============================== consumer warps ==============================
    T acc[VPT][V];
    int slot = 0;
    uint32_t par = 0;
    const unsigned char* slot_p = smem + kHdrAligned + kPad + tid * 16;  // this thread's vector 0, stream 0
    while (true) {
      mbar_wait(full0 + 8 * slot, par);
      const FastMeta& M = meta[slot];
      const int flags = M.flags;
      if (flags & F_END) break;
      const int nvalid = M.nvalid;
      const int nterms = M.nterms;
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
      for (int t = 0; t < nterms; ++t) {
        const GTerm gt = M.terms[t];
        FastIO<T> io;
        io.stride = kBufBytes;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int e0 = (i * kConsumers + tid) * V;
          const bool first = (flags & F_BLK0) && e0 == 0;
          const int last = (flags & F_BLKEND) ? nvalid - 1 - e0 : (1 << 30);
          T val[V];
          io.b0 = reinterpret_cast<const char*>(slot_p) + gt.stream0 * kBufBytes + i * kConsumers * 16;
          eval_fast<T>(gt.pattern, io, stages + gt.stage0, first, last, val);
          if (gt.sign >= 0) {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[i][j] = acc[i][j] + val[j];
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j) acc[i][j] = acc[i][j] - val[j];
          }
        }
      }
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * slot);  // slot may be refilled
      if (flags & F_LAST) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
          const int e0 = (i * kConsumers + tid) * V;
          if (e0 + V <= nvalid) {
            Vec v;
            T* vs = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int j = 0; j < V; ++j) vs[j] = acc[i][j];
            *reinterpret_cast<Vec*>(out_tile + e0) = v;
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j)
              if (e0 + j < nvalid) out_tile[e0 + j] = acc[i][j];
          }
        }
      }
      slot_p += slot_bytes;
      if (++slot == nslots) { slot = 0; par ^= 1; slot_p -= (size_t)nslots * slot_bytes; }
    }
  }
}


template <typename T, int CW, int VPT>
void launch_variant(const DevFused& f, FastParams& P, cudaStream_t s) {
  constexpr int kBufBytes = CW * 32 * 16 * VPT + 2 * kPad;
  int n = (kSmemLimit - kHdrAligned) / (f.slot_streams * kBufBytes);
  if (n > kMaxSlots) n = kMaxSlots;
  P.nslots = n;
  const size_t smem = kHdrAligned + (size_t)n * f.slot_streams * kBufBytes;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(jets_fused_fast_kernel<T, CW, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  const int64_t Q = P.nitems * P.S;
  int64_t grid = ctx().sm_count;
  if (grid > Q) grid = Q > 0 ? Q : 1;
  jets_fused_fast_kernel<T, CW, VPT><<<(unsigned)grid, CW * 32 + 32, smem, s>>>(P);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

template <typename T>
void launch_dtype(const DevFused& f, FastParams& P, cudaStream_t s) {
  switch (f.variant) {
    case 1: launch_variant<T, 8, 2>(f, P, s); break;
    case 2: launch_variant<T, 16, 2>(f, P, s); break;
    default: launch_variant<T, 16, 1>(f, P, s); break;
  }
}

}  // namespace

int fast_tile_bytes(int variant) { return variant == 2 ? 16384 : 8192; }
int fast_nslots(int variant, int slot_streams) {
  const int n = (kSmemLimit - kHdrAligned) / (slot_streams * (fast_tile_bytes(variant) + 2 * kPad));
  return n > kMaxSlots ? kMaxSlots : n;
}

void launch_fused_fast(const DevFused& f, int dtype, const char* in, char* out, cudaStream_t s) {
  if (f.nrows == 0 || f.ntiles == 0) return;
  FastParams P;
  P.groups = f.groups; P.rows = f.rows; P.segs = f.segs; P.order = f.order;
  P.nsegs = f.nsegs; P.nslots = 0; P.slot_streams = f.slot_streams; P.S = f.S;
  P.nitems = f.nitems; P.in = in; P.out = out; P.hl = f.hl; P.hr = f.hr;
  if (dtype == JETS_F32) launch_dtype<float>(f, P, s);
  else launch_dtype<double>(f, P, s);
}

}  // namespace jets
