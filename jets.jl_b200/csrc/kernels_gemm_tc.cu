// Dense block primitive applied to MANY right-hand sides: D = A X and M = A' Y for a whole table
// of column-major Float32 matrix blocks on the 5th-generation tensor cores
// (replaces the per-column stdlib mul! the reference would run for a matrix of right-hand sides:
// _matmul_df!/_df'! src/Jets.jl:573-574 and the JetBlock_df!/df'! accumulation :1024,:1049).
//
// Precision: kind::tf32 keeps 11 significant bits, far short of the 1e-5 the path promises, so
// every product is evaluated as the classic three-term split
//     a*x ~= a_hi*x_hi + a_hi*x_lo + a_lo*x_hi,   a_hi = a & ~0x1fff,  a_lo = tf32(a - a_hi)
// (all four operands exactly representable in tf32, so the tensor core sees exact inputs; the
// dropped a_lo*x_lo term is 2^-22 relative).  The tensor core TRUNCATES when it adds into the f32
// accumulator (measured: 1.5e-5 relative after 768 accumulates), so the main term a_hi*x_hi gets
// its own accumulator columns (one truncation per 8-deep k-step), the two 2^-11-sized correction
// terms share a second set, and both are folded into per-thread f32 registers with
// round-to-nearest adds every kDrainStages k-stages: the result does not depend on the length
// of the reduction.
// Default ("mixed") mode: the two correction terms run on kind::f16 with bf16 operands, K-concatenated
// into one chain -- bf(a_lo)*bf(x) + bf(a)*bf(x_lo), round-to-nearest conversions, 2^-19 of the
// product per term and zero-mean -- which takes 256 tensor-pipe cycles per k-stage instead of 384
// and a third less operand traffic out of shared memory; measured error 1.3e-6 against Float64
// (all-tf32: 1.6e-6).  JETS_B200_TC_MIXED=0 keeps all three terms on tf32.
//
// The kernel is HBM-bound by design (32 flop/B on the matrix stream), so it is built around the
// matrix bytes: one CTA per SM, persistent over 128-row output tiles, no split-K, no atomics.
//   warp 8      A producer: per 32-deep k-stage four 32x32 TMA boxes of the matrix block
//               (cp.async.bulk.tensor, evict-first), an 8-deep ring = 128 KB in flight per SM
//   warp 9      X producer: one box of the pre-split right-hand sides [x_hi | x_lo] per stage
//               (evict-last: every tile re-reads them from L2), a 4-deep ring
//   warps 4-7, 12-15   read the landed A tile ONCE from shared memory, one matrix row (half a stage's k) per thread, split it
//               and write a_hi / a_lo straight into TENSOR MEMORY (tcgen05.st): the MMA takes its A
//               operand from TMEM, so shared memory carries the matrix bytes exactly twice
//               (TMA write + this read) instead of ~8x with hi/lo copies in shared memory
//   warps 10,11 issue the MMAs (A from TMEM, X from shared memory through a K-major 128B-swizzle
//               descriptor): warp 10 the tf32 main term, warp 11 the bf16 corrections -- disjoint
//               accumulator columns, so the two chains need no ordering between them; both commit
//               ring slots and accumulator chunks to mbarriers.  Everything the issuers touch is
//               warp-uniform BY CONSTRUCTION (warp index and TMEM base broadcast with shfl), so the
//               compiler keeps it in uniform registers and emits the UTCHMMAs back to back; with a
//               lane-derived warp index it wrapped every MMA and commit in an "elect one of the
//               remaining threads" loop and the single issuing lane (~150 dependent instructions,
//               ~1100 cycles per k-stage) was the kernel's bottleneck: 1.15 ms -> 0.74 ms on 4 GiB.
//   warps 0-3   epilogue: tcgen05.ld the accumulator chunk, add into registers, store the tile
// Both orientations read the matrix from its one column-major layout: A*X stages [k][32 rows]
// boxes and each thread gathers its row with conflict-free 4-byte loads; A'*Y stages [col][32 k]
// boxes (128B swizzle) and each thread reads its own 128-byte row.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <map>
#include "common.hpp"

namespace jets {
namespace {

constexpr int BM = 128;            // output rows per tile (= TMEM lanes)
constexpr int BK = 32;             // k per stage
// Ring depths are template parameters of the kernel (kAStages matrix stages, kXStages right-hand-side
// stages; both 16 KB).  Measured: 6 + 8 is 9% slower than 8 + 4 -- the matrix ring is what hides HBM latency.
constexpr int kTStages = 4;        // TMEM operand ring (hi|lo slots of 64 columns), own barriers
constexpr int kABytes = BM * BK * 4;           // 16 KB
constexpr int kXBytes = 128 * BK * 4;          // 2*NP rows of 128 B, NP <= 64
constexpr int kDrainStages = 8;    // TMEM accumulation length (256 k) before folding into registers
constexpr int kThreads = 512;      // 16 warps: 0-3 epilogue, 4-7 + 12-15 split, 8 A producer, 9 X producer, 10 MMA
constexpr int kTmemCols = 512;     // [0,256): 2 accumulator buffers x (main | correction) x 64; [256,512): kTStages A slots x (hi | lo) x 32
constexpr int kBufCols = 128;
constexpr int kASlotCol = 256;
constexpr int smem_bytes(int a_stages, int x_stages) { return 1024 /*align slack*/ + a_stages * kABytes + x_stages * kXBytes + 1024; }

struct TcParams {
  const DBlock* blocks;
  const int32_t* row_ptr;
  const int32_t* tile_ptr;
  const int32_t* koff;        // per block: first k of its input block inside the split buffer
  const CUtensorMap* amaps;   // per block
  int32_t ngroups, ntiles;
  int32_t np;                 // padded right-hand-side count (multiple of 16, <= 64)
  int32_t nrhs, n0;           // right-hand sides handled by this launch: [n0, n0+nrhs)
  int32_t trans, acc;
  float* out;
  int32_t expt;               // JETS_B200_TC_EXPT: 1 = skip the a_lo MMAs, 2 = skip all MMAs (timing experiments only)
  int32_t mixed;              // 1: correction terms on kind::f16 (bf16 operands), see umma_stage_corr_bf16; 0: all three terms tf32
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
__device__ __forceinline__ void tma_2d(uint32_t dst, const void* map, int c0, int c1, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, A: 128 lanes x 8 columns of tf32, B: N x 8 through a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One k-stage (BK = 32 = four tf32 k-steps) of the split product in ONE asm block:
//   D_main|corr (+)= a_hi x [x_hi | x_lo]   and   D_corr += a_lo x x_hi   for each k-step.
// A k-step advances the TMEM operand by 8 columns and the shared-memory descriptor by 32 bytes
// (start-address field += 2).  The issuing thread is the pipeline's narrowest point (ncu: ~155
// single-lane instructions per stage with one asm block per MMA, ~900 cycles, while the tensor
// pipe needs 384), so everything loop-invariant is folded into this block's immediates.
__device__ __forceinline__ void umma_stage_tf32(uint32_t d_main, uint32_t d_corr, uint32_t a_hi, uint32_t a_lo, uint64_t xdesc,
                                                uint32_t idesc_main, uint32_t idesc_corr, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .b64 x1, x2, x3;\n\t"
      ".reg .b32 h1, h2, h3, l1, l2, l3;\n\t"
      "setp.ne.b32 p0, %7, 0;\n\t"
      "setp.ne.b32 p1, %5, 0;\n\t"          // always true (an instruction descriptor is never 0)
      "add.s64 x1, %4, 2;\n\t"
      "add.s64 x2, %4, 4;\n\t"
      "add.s64 x3, %4, 6;\n\t"
      "add.u32 h1, %2, 8;\n\t"
      "add.u32 h2, %2, 16;\n\t"
      "add.u32 h3, %2, 24;\n\t"
      "add.u32 l1, %3, 8;\n\t"
      "add.u32 l2, %3, 16;\n\t"
      "add.u32 l3, %3, 24;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %4, %5, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], [%3], %4, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [h1], x1, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], [l1], x1, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [h2], x2, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], [l2], x2, %6, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [h3], x3, %5, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%1], [l3], x3, %6, p1;\n\t"
      "}" ::"r"(d_main), "r"(d_corr), "r"(a_hi), "r"(a_lo), "l"(xdesc), "r"(idesc_main), "r"(idesc_corr), "r"(accumulate)
      : "memory");
}
// The same k-stage with the two correction terms evaluated on bf16 operands (kind::f16, K = 16 per
// MMA): a*x = a_hi*x_hi + bf(a)*bf(x_lo) + bf(a_lo)*bf(x) + O(2^-19 |a x|) with round-to-nearest bf16
// conversions (zero-mean errors of <= 2^-9 on terms that are <= 2^-10 of the product).  The main term
// stays tf32 with N = np; the corrections are 4 bf16 MMAs of N = np instead of the tf32 path's
// 4 x (N = np extra columns) + 4 x (N = np): 256 tensor-pipe cycles per stage instead of 384, and a
// third less operand traffic out of shared memory.
//   TMEM operand slot (64 columns): [0,32) a_hi tf32 k 0..31 | [32,40) bf(a_lo) k 0..15 | [40,48) bf(a) k 0..15
//                                   | [48,56) bf(a_lo) k 16..31 | [56,64) bf(a) k 16..31     (two bf16 per column)
//   X slot rows [np,2np) (128 B each): bf(x) k 0..31 | bf(x_lo) k 0..31
__device__ __forceinline__ void umma_stage_main_tf32(uint32_t d_main, uint32_t a_slot, uint64_t xdesc, uint32_t idesc_tf32,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .b64 x1, x2, x3;\n\t"
      ".reg .b32 h1, h2, h3;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\t"
      "setp.ne.b32 p1, %3, 0;\n\t"          // always true
      "add.s64 x1, %2, 2;\n\t"
      "add.s64 x2, %2, 4;\n\t"
      "add.s64 x3, %2, 6;\n\t"
      "add.u32 h1, %1, 8;\n\t"
      "add.u32 h2, %1, 16;\n\t"
      "add.u32 h3, %1, 24;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [h1], x1, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [h2], x2, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [h3], x3, %3, p1;\n\t"
      "}" ::"r"(d_main), "r"(a_slot), "l"(xdesc), "r"(idesc_tf32), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_stage_corr_bf16(uint32_t d_corr, uint32_t a_slot, uint64_t xbdesc, uint32_t idesc_bf16,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p0, p1;\n\t"
      ".reg .b64 b1, b2, b3;\n\t"
      ".reg .b32 c0, c1, c2, c3;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\t"
      "setp.ne.b32 p1, %3, 0;\n\t"          // always true
      "add.s64 b1, %2, 4;\n\t"              // bf(x_lo) k 0..15   (+64 B)
      "add.s64 b2, %2, 2;\n\t"              // bf(x)    k 16..31  (+32 B)
      "add.s64 b3, %2, 6;\n\t"              // bf(x_lo) k 16..31  (+96 B)
      "add.u32 c0, %1, 32;\n\t"
      "add.u32 c1, %1, 40;\n\t"
      "add.u32 c2, %1, 48;\n\t"
      "add.u32 c3, %1, 56;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [c0], %2, %3, p0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [c1], b1, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [c2], b2, %3, p1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [c3], b3, %3, p1;\n\t"
      "}" ::"r"(d_corr), "r"(a_slot), "l"(xbdesc), "r"(idesc_bf16), "r"(accumulate)
      : "memory");
}
// One lane of a fully converged warp (CUTLASS's elect_one_sync): the compiler knows the branch it
// guards runs in a single thread and predicates the uniform-datapath instructions directly.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
// two floats -> packed bf16x2, round to nearest even; `lo` lands in bits [0,16) (the lower k index)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31]) : "memory");
}

// Shared-memory matrix descriptor (sm_100 UMMA), K-major 128B swizzle: start>>4 [0,14), LBO>>4 [16,30)
// (unused for swizzled K-major), SBO>>4 [32,46) = 1024 B between 8-row groups, version=1 [46,48),
// layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2, both K-major, N>>3 [17,23), M>>4 [24,29).
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// The same for kind::f16 with bf16 operands: A = B = BF16 (format code 1).
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ int find_group(const int32_t* tile_ptr, int ngroups, int tile) {
  int lo = 0, hi = ngroups - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_ptr[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int kAStages, int kXStages>
__global__ void __launch_bounds__(kThreads, 1)
jets_gemm_tc_kernel(const TcParams P, const __grid_constant__ CUtensorMap xmap) {
  static_assert(kAStages <= 16 && kXStages <= 16, "barrier arrays hold 16 stages");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // 128B-swizzle atoms need 1024-byte alignment
  unsigned char* sbase = smem_raw + (base - raw);
  const uint32_t a0 = base, x0 = base + kAStages * kABytes;
  const uint32_t bars = x0 + kXStages * kXBytes;
  const uint32_t a_full0 = bars, a_empty0 = bars + 128, x_full0 = bars + 256, x_empty0 = bars + 384, t_ready0 = bars + 512,
                 mma_done0 = bars + 544, acc_full0 = bars + 576, acc_empty0 = bars + 592;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sbase + kAStages * kABytes + kXStages * kXBytes + 640);
  // warp-uniform by construction (shuffle from lane 0), so that everything the MMA issuers derive from it can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(a_full0 + 8 * s, 1);
      mbar_init(a_empty0 + 8 * s, 8);
    }
    for (int s = 0; s < kXStages; ++s) {
      mbar_init(x_full0 + 8 * s, 1);
      mbar_init(x_empty0 + 8 * s, 2);     // both MMA issuers commit
    }
    for (int s = 0; s < kTStages; ++s) {
      mbar_init(t_ready0 + 8 * s, 8);
      mbar_init(mma_done0 + 8 * s, 2);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full0 + 8 * b, 2);
      mbar_init(acc_empty0 + 8 * b, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 10) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  const int np = P.np;

  if (warp == 8) {
    // =============================== A producer =================================
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int g = find_group(P.tile_ptr, P.ngroups, tile);
        const int m0 = (tile - P.tile_ptr[g]) * BM;
        for (int e = P.row_ptr[g]; e < P.row_ptr[g + 1]; ++e) {
          const DBlock b = P.blocks[e];
          const int kdim = P.trans ? b.rows : b.cols;
          const int nk = (kdim + BK - 1) / BK;
          const CUtensorMap* amap = P.amaps + e;
          for (int kc = 0; kc < nk; ++kc, ++it) {
            const int s = it % kAStages;
            mbar_wait(a_empty0 + 8 * s, ((it / kAStages) & 1) ^ 1);
            const uint32_t st = a0 + s * kABytes;
            mbar_expect_tx(a_full0 + 8 * s, kABytes);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              // tensor map dims: (row, col) -- forward: rows are the tile's M, cols the k-stage;
              // adjoint: rows are the k-stage, cols the tile's M
              const int c_row = P.trans ? kc * BK : m0 + 32 * q;
              const int c_col = P.trans ? m0 + 32 * q : kc * BK;
              tma_2d(st + q * 4096, amap, c_row, c_col, a_full0 + 8 * s, pol);
            }
          }
        }
      }
    }
  } else if (warp == 9) {
    // =============================== X producer =================================
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      const uint32_t xbytes = (uint32_t)(2 * np) * BK * 4;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const int g = find_group(P.tile_ptr, P.ngroups, tile);
        for (int e = P.row_ptr[g]; e < P.row_ptr[g + 1]; ++e) {
          const DBlock b = P.blocks[e];
          const int kdim = P.trans ? b.rows : b.cols;
          const int nk = (kdim + BK - 1) / BK;
          const int koff = P.koff[e];
          for (int kc = 0; kc < nk; ++kc, ++it) {
            const int s = it % kXStages;
            mbar_wait(x_empty0 + 8 * s, ((it / kXStages) & 1) ^ 1);
            mbar_expect_tx(x_full0 + 8 * s, xbytes);
            tma_2d(x0 + s * kXBytes, &xmap, koff + kc * BK, 0, x_full0 + 8 * s, pol);
          }
        }
      }
    }
  } else if (warp == 10 || warp == 11) {
    // =============================== MMA issuers ================================
    // Two warps run the same control flow and issue disjoint accumulator chains: warp 10 the main
    // term (columns [0,np) of the buffer; in the all-tf32 mode the whole stage), warp 11 the bf16
    // corrections (columns [np,2np)).  ncu showed ONE issuing lane as the kernel's narrowest point
    // (~150 dependent single-lane instructions per k-stage, ~1100 cycles, against 256 tensor-pipe
    // cycles and ~700 cycles of HBM time); two issuers, an elect.sync leader (no per-instruction
    // "any thread left?" loop around the uniform-datapath MMAs) and hoisted descriptors cut that.
    // Ring slots and accumulator chunks are committed by both warps (their barriers count 2).
    const bool corr_warp = warp == 11;
    const uint32_t idesc_main = make_idesc(2 * np);   // a_hi x [x_hi | x_lo]
    const uint32_t idesc_corr = make_idesc(np);       // a_lo x x_hi, or the N = np main term of the mixed mode
    const uint32_t idesc_bf16 = make_idesc_bf16(np);  // mixed mode: bf16 corrections
    const uint64_t xdesc0 = make_desc_k128(x0);       // X slot s: start-address field += s * kXBytes / 16
    const uint64_t xb_off = (uint64_t)((np * 128) >> 4);   // rows [np, 2np) of an X slot
    const bool leader = elect_one();
    const int mixed = P.mixed, expt = P.expt;
    uint32_t it = 0, cn = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const int g = find_group(P.tile_ptr, P.ngroups, tile);
      for (int e = P.row_ptr[g]; e < P.row_ptr[g + 1]; ++e) {
        const DBlock b = P.blocks[e];
        const int kdim = P.trans ? b.rows : b.cols;
        const int nk = (kdim + BK - 1) / BK;
        for (int c0 = 0; c0 < nk; c0 += kDrainStages, ++cn) {
          const int c1 = min(nk, c0 + kDrainStages);
          const int buf = cn & 1;
          mbar_wait(acc_empty0 + 8 * buf, ((cn >> 1) & 1) ^ 1);
          // columns [0,np): a_hi*x_hi (the main term, ONE truncating accumulate per k-step);
          // columns [np,2np): the two correction terms (2^-11 smaller, so their rounding does not matter)
          const uint32_t d_tmem = tmem_base + buf * kBufCols;
          for (int kc = c0; kc < c1; ++kc, ++it) {
            const int s = it % kXStages, t = it % kTStages;
            mbar_wait(x_full0 + 8 * s, (it / kXStages) & 1);
            mbar_wait(t_ready0 + 8 * t, (it / kTStages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (leader) {
              const uint32_t a_hi = tmem_base + kASlotCol + t * 64;
              const uint64_t xd = xdesc0 + (uint64_t)(s * (kXBytes >> 4));
              const uint32_t accf = kc > c0 ? 1u : 0u;
              if (expt == 0) {
                if (mixed) {
                  if (!corr_warp) umma_stage_main_tf32(d_tmem, a_hi, xd, idesc_corr, accf);
                  else umma_stage_corr_bf16(d_tmem + np, a_hi, xd + xb_off, idesc_bf16, accf);
                } else if (!corr_warp) {
                  umma_stage_tf32(d_tmem, d_tmem + np, a_hi, a_hi + 32, xd, idesc_main, idesc_corr, accf);
                }
              } else if (expt == 3 && !corr_warp) {   // timing experiment: only the four N = np tf32 main MMAs
                umma_stage_main_tf32(d_tmem, a_hi, xd, idesc_corr, accf);
              } else if (expt == 1 && !corr_warp) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma_tf32_ts(d_tmem, a_hi + ks * 8, xd + (uint64_t)(2 * ks), idesc_main, (kc > c0 || ks > 0) ? 1u : 0u);
              }
              umma_commit(x_empty0 + 8 * s);                       // right-hand-side slot free once the MMAs retire
              umma_commit(mma_done0 + 8 * t);                      // TMEM operand slot likewise
              if (kc + 1 == c1) umma_commit(acc_full0 + 8 * buf);  // accumulator chunk complete
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // =============================== split A into TMEM (warps 4-7 and 12-15) =====
    // Two warps per TMEM lane quarter: warps 4-7 take k columns [0,16) of the stage, warps 12-15
    // columns [16,32) -- one warp per quarter could not split a stage as fast as HBM delivers it.
    const int q = warp & 3;                 // TMEM lane quarter == 32-row box of the tile
    const int half = warp >= 12 ? 1 : 0;    // which 16 of the stage's 32 k columns
    const int r = q * 32 + lane;            // this thread's tile row
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const int g = find_group(P.tile_ptr, P.ngroups, tile);
      for (int e = P.row_ptr[g]; e < P.row_ptr[g + 1]; ++e) {
        const int kdim = P.trans ? P.blocks[e].rows : P.blocks[e].cols;
        const int nk = (kdim + BK - 1) / BK;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int sa = it % kAStages, st = it % kTStages;
          mbar_wait(a_full0 + 8 * sa, (it / kAStages) & 1);
          uint32_t hi[16], lo[16];
          const unsigned char* tile_p = sbase + (size_t)sa * kABytes;
          if (P.trans) {
            // [col][32 k], 128B swizzle: 16-byte chunk c of row r sits at chunk (c ^ (r & 7))
            const uint4* row = reinterpret_cast<const uint4*>(tile_p + r * 128);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 v = row[(4 * half + c) ^ (r & 7)];
              hi[4 * c + 0] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
            }
          } else {
            // [k][32 rows] per 32-row box, no swizzle: the warp reads one 128-byte line per k
            const uint32_t* col = reinterpret_cast<const uint32_t*>(tile_p + q * 4096) + lane + half * 16 * 32;
#pragma unroll
            for (int k = 0; k < 16; ++k) hi[k] = col[k * 32];
          }
          if (P.mixed) {
            // lo[0,8) = bf(a_lo) pairs, lo[8,16) = bf(a) pairs of this thread's 16 k
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
              const float a0 = __uint_as_float(hi[k]), a1 = __uint_as_float(hi[k + 1]);
              const uint32_t h0 = hi[k] & 0xFFFFE000u, h1 = hi[k + 1] & 0xFFFFE000u;
              lo[k >> 1] = pack_bf16x2(a0 - __uint_as_float(h0), a1 - __uint_as_float(h1));
              lo[8 + (k >> 1)] = pack_bf16x2(a0, a1);
              hi[k] = h0; hi[k + 1] = h1;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const uint32_t a = hi[k];
              hi[k] = a & 0xFFFFE000u;
              lo[k] = __float_as_uint(__uint_as_float(a) - __uint_as_float(hi[k])) & 0xFFFFE000u;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(a_empty0 + 8 * sa);            // the shared-memory slot can be refilled
          mbar_wait(mma_done0 + 8 * st, ((it / kTStages) & 1) ^ 1);  // TMEM operand slot free (MMAs of it-4 retired)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem_base + lane_addr + kASlotCol + st * 64 + half * 16;
          tmem_st16(taddr, hi);
          tmem_st16(taddr + 32, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(t_ready0 + 8 * st);
        }
      }
    }
  } else {
    // =============================== epilogue (warps 0-3) =======================
    float acc[64];
    uint32_t cn = 0;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
      const int g = find_group(P.tile_ptr, P.ngroups, tile);
      const int m0 = (tile - P.tile_ptr[g]) * BM;
#pragma unroll
      for (int c = 0; c < 64; ++c) acc[c] = 0.f;
      int64_t out_off = 0;
      int mdim = 0;
      for (int e = P.row_ptr[g]; e < P.row_ptr[g + 1]; ++e) {
        const DBlock b = P.blocks[e];
        out_off = b.out_off;
        mdim = P.trans ? b.cols : b.rows;
        const int kdim = P.trans ? b.rows : b.cols;
        const int nk = (kdim + BK - 1) / BK;
        for (int c0 = 0; c0 < nk; c0 += kDrainStages, ++cn) {
          const int buf = cn & 1;
          mbar_wait(acc_full0 + 8 * buf, (cn >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem_base + lane_base + buf * kBufCols;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q * 16 < np) {
              uint32_t v[16], w[16];
              tmem_ld16(taddr + q * 16, v);
              tmem_ld16(taddr + np + q * 16, w);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int c = 0; c < 16; ++c) acc[q * 16 + c] += __uint_as_float(v[c]) + __uint_as_float(w[c]);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty0 + 8 * buf);
        }
      }
      // store the tile: row i of the block, right-hand side n -> out[out_off + i + n*mdim]
      const int i = m0 + warp * 32 + lane;
      if (i < mdim) {
        float* o = P.out + out_off + i + (int64_t)P.n0 * mdim;
#pragma unroll
        for (int n = 0; n < 64; ++n) {
          if (n < P.nrhs) {
            float* p = o + (int64_t)n * mdim;
            if (P.acc == ACC_SET) *p = acc[n];
            else if (P.acc == ACC_ADD) *p = *p + acc[n];
            else *p = *p - acc[n];
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 10) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// Right-hand sides -> the split buffer: xs[n][koff + k] = hi(x[k + n*kdim]), xs[np + n][...] = lo.
struct SplitSeg { int64_t in_off; int32_t kdim, koff; };
__global__ void split_rhs_kernel(const float* __restrict__ in, float* __restrict__ xs, const SplitSeg* __restrict__ segs,
                                 int nsegs, int64_t kp, int np, int nrhs, int n0, int nrhs_total, int mixed) {
  const int seg = blockIdx.y;
  const SplitSeg sg = segs[seg];
  const int64_t total = (int64_t)sg.kdim * nrhs;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / sg.kdim);
    const int k = (int)(idx - (int64_t)n * sg.kdim);
    const float x = in[sg.in_off + k + (int64_t)(n0 + n) * sg.kdim];
    const uint32_t h = __float_as_uint(x) & 0xFFFFE000u;
    const uint32_t l = __float_as_uint(x - __uint_as_float(h)) & 0xFFFFE000u;
    xs[(int64_t)n * kp + sg.koff + k] = __uint_as_float(h);
    if (mixed) {
      // row np+n holds, per 32-deep k-stage (128 bytes): bf(x) k 0..31 | bf(x_lo) k 0..31
      __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(xs + (int64_t)(np + n) * kp + sg.koff + (k & ~31));
      row[k & 31] = __float2bfloat16_rn(x);
      row[32 + (k & 31)] = __float2bfloat16_rn(x - __uint_as_float(h));
    } else {
      xs[(int64_t)(np + n) * kp + sg.koff + k] = __uint_as_float(l);
    }
  }
  (void)nsegs; (void)nrhs_total;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    JETS_CHECK(p && q == cudaDriverEntryPointSuccess, JETS_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
void encode_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t stride_bytes, uint32_t box_inner,
               uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t es[2] = {1, 1};
  const CUresult r = encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JETS_CHECK(r == CUDA_SUCCESS, JETS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p dims=%llu x %llu stride=%llu", (int)r, ptr,
             (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)stride_bytes);
}

}  // namespace

// Can this table of dense blocks run on the tensor-core path?
bool gemm_tc_eligible(int dtype, const std::vector<DBlock>& blocks) {
  if (dtype != JETS_F32 || blocks.empty()) return false;
  if (getenv("JETS_B200_NO_TC")) return false;
  for (const DBlock& b : blocks) {
    if (b.nrhs < 2 || b.nrhs != blocks[0].nrhs) return false;
    if ((reinterpret_cast<uintptr_t>(b.A) & 15) || (b.lda % 4)) return false;  // TMA: 16-byte base and row pitch
  }
  return true;
}

// Builds the device tables of a tensor-core GEMM step (blocks sorted by out_off, block-level offsets).
void gemm_tc_prepare(Step& st, Plan& plan) {
  const std::vector<DBlock>& blocks = st.dblocks;
  const bool trans = blocks[0].trans != 0;
  const int nrhs = blocks[0].nrhs;
  const int ngrp = std::min(nrhs, 64);
  const int np = (ngrp + 15) & ~15;
  std::vector<int32_t> row_ptr{0}, tile_ptr{0}, koff(blocks.size());
  std::map<int64_t, int32_t> seg_of;      // in_off -> koff
  std::vector<SplitSeg> segs;
  int64_t kp = 0;
  for (size_t i = 0; i < blocks.size(); ++i) {
    const DBlock& b = blocks[i];
    const int kdim = trans ? b.rows : b.cols;
    auto it = seg_of.find(b.in_off);
    if (it == seg_of.end()) {
      JETS_CHECK(kp + kdim + BK < (1LL << 31), JETS_ERR_UNSUPPORTED, "dense block table too large for the tensor-core path");
      it = seg_of.emplace(b.in_off, (int32_t)kp).first;
      segs.push_back(SplitSeg{b.in_off, kdim, (int32_t)kp});
      kp += ((int64_t)kdim + BK - 1) / BK * BK;   // zero padding up to the next k-stage
    }
    koff[i] = it->second;
    const bool last = i + 1 == blocks.size() || blocks[i + 1].out_off != b.out_off;
    if (last) {
      row_ptr.push_back((int32_t)i + 1);
      const int mdim = trans ? b.cols : b.rows;
      tile_ptr.push_back(tile_ptr.back() + (mdim + BM - 1) / BM);
    }
  }
  st.n_out_rows = (int32_t)row_ptr.size() - 1;
  st.gemv_tiles = tile_ptr.back();
  st.tc_np = np;
  st.tc_kp = kp;
  st.tc_nsegs = (int32_t)segs.size();
  // split buffer (zero-filled once: the padding is never written afterwards)
  const size_t xs_bytes = (size_t)2 * np * kp * sizeof(float);
  char* xs = nullptr;
  CUDA_TRY(cudaMalloc(&xs, xs_bytes));
  CUDA_TRY(cudaMemset(xs, 0, xs_bytes));
  plan.blobs.push_back(xs);
  st.tc_xs = xs;
  encode_2d(reinterpret_cast<CUtensorMap*>(st.tc_xmap), xs, (uint64_t)kp, (uint64_t)2 * np, (uint64_t)kp * 4, BK, 2 * np);
  // per-block matrix maps: dims (rows, cols), row pitch lda
  std::vector<CUtensorMap> maps(blocks.size());
  for (size_t i = 0; i < blocks.size(); ++i)
    encode_2d(&maps[i], blocks[i].A, (uint64_t)blocks[i].rows, (uint64_t)blocks[i].cols, (uint64_t)blocks[i].lda * 4, 32, 32,
              trans ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_maps = 0, o_blocks = o_maps + al(maps.size() * sizeof(CUtensorMap)),
               o_row = o_blocks + al(blocks.size() * sizeof(DBlock)), o_tile = o_row + al(row_ptr.size() * 4),
               o_koff = o_tile + al(tile_ptr.size() * 4), o_segs = o_koff + al(koff.size() * 4),
               total = o_segs + al(segs.size() * sizeof(SplitSeg));
  std::vector<char> host(total, 0);
  memcpy(host.data() + o_maps, maps.data(), maps.size() * sizeof(CUtensorMap));
  memcpy(host.data() + o_blocks, blocks.data(), blocks.size() * sizeof(DBlock));
  memcpy(host.data() + o_row, row_ptr.data(), row_ptr.size() * 4);
  memcpy(host.data() + o_tile, tile_ptr.data(), tile_ptr.size() * 4);
  memcpy(host.data() + o_koff, koff.data(), koff.size() * 4);
  memcpy(host.data() + o_segs, segs.data(), segs.size() * sizeof(SplitSeg));
  char* blob = nullptr;
  CUDA_TRY(cudaMalloc(&blob, total));
  CUDA_TRY(cudaMemcpy(blob, host.data(), total, cudaMemcpyHostToDevice));
  plan.blobs.push_back(blob);
  st.tc_maps = blob + o_maps;
  st.d_dblocks = reinterpret_cast<DBlock*>(blob + o_blocks);
  st.d_row_ptr = reinterpret_cast<int32_t*>(blob + o_row);
  st.tc_tile_ptr = reinterpret_cast<int32_t*>(blob + o_tile);
  st.tc_koff = reinterpret_cast<int32_t*>(blob + o_koff);
  st.tc_segs = blob + o_segs;
}

// JETS_B200_TC_MIXED=0 keeps all three split terms on kind::tf32 (the A/B baseline); default: bf16 corrections.
static int tc_mixed() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("JETS_B200_TC_MIXED"); v = e ? (atoi(e) != 0) : 1; }
  return v;
}

// JETS_B200_TC_RING selects the (matrix, right-hand-side) ring depths: 0 = 8+4, 1 = 10+4, 2 = 11+3, 3 = 9+5, 4 = 12+2.
static int tc_ring() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("JETS_B200_TC_RING"); v = e ? atoi(e) : 0; }
  return v;
}
template <int AS, int XS>
static void launch_tc(int grid, const TcParams& P, const CUtensorMap& xmap, cudaStream_t s) {
  constexpr int bytes = smem_bytes(AS, XS);
  static_assert(bytes <= 227 * 1024, "ring does not fit shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(jets_gemm_tc_kernel<AS, XS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    attr_set = true;
  }
  jets_gemm_tc_kernel<AS, XS><<<grid, kThreads, bytes, s>>>(P, xmap);
}

void launch_gemm_tc(const Step& st, const char* in, char* out, cudaStream_t s) {
  if (st.n_out_rows == 0 || st.gemv_tiles == 0) return;
  const int ring = tc_ring();
  const int nrhs_total = st.dblocks[0].nrhs;
  int maxk = 1;
  for (const DBlock& b : st.dblocks) maxk = std::max(maxk, st.dblocks[0].trans ? b.rows : b.cols);
  for (int n0 = 0; n0 < nrhs_total; n0 += 64) {
    const int nrhs = std::min(64, nrhs_total - n0);
    {
      const int64_t work = (int64_t)maxk * nrhs;
      dim3 grid((unsigned)std::min<int64_t>((work + 255) / 256, 64), (unsigned)st.tc_nsegs);
      split_rhs_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(st.tc_xs),
                                            reinterpret_cast<const SplitSeg*>(st.tc_segs), st.tc_nsegs, st.tc_kp, st.tc_np, nrhs, n0,
                                            nrhs_total, tc_mixed());
      CUDA_TRY(cudaGetLastError());
      count_launch();
    }
    TcParams P;
    P.blocks = st.d_dblocks;
    P.row_ptr = st.d_row_ptr;
    P.tile_ptr = st.tc_tile_ptr;
    P.koff = st.tc_koff;
    P.amaps = reinterpret_cast<const CUtensorMap*>(st.tc_maps);
    P.ngroups = st.n_out_rows;
    P.ntiles = (int32_t)st.gemv_tiles;
    P.np = st.tc_np;
    P.nrhs = nrhs;
    P.n0 = n0;
    P.trans = st.dblocks[0].trans;
    P.acc = st.acc;
    P.out = reinterpret_cast<float*>(out);
    { const char* e = getenv("JETS_B200_TC_EXPT"); P.expt = e ? atoi(e) : 0; }
    P.mixed = tc_mixed();
    const int grid = (int)std::min<int64_t>(st.gemv_tiles, ctx().sm_count);
    CUtensorMap xmap;
    memcpy(&xmap, st.tc_xmap, sizeof(xmap));
    switch (ring) {
      case 1: launch_tc<10, 4>(grid, P, xmap, s); break;
      case 2: launch_tc<11, 3>(grid, P, xmap, s); break;
      case 3: launch_tc<9, 5>(grid, P, xmap, s); break;
      case 4: launch_tc<12, 2>(grid, P, xmap, s); break;
      default: launch_tc<8, 4>(grid, P, xmap, s); break;
    }
    CUDA_TRY(cudaGetLastError());
    count_launch();
  }
}

}  // namespace jets
