// Dense block primitive: d = A m and m = A' d on column-major matrices (Julia layout), for a
// whole table of blocks in one launch (replaces _matmul_df!/_df'! src/Jets.jl:573-574 and the
// per-block accumulate of JetBlock_df!/df'! :1024,:1049 for dense leaves; fixture JopBaz
// test/runtests.jl:27-33).  Single-vector GEMV is HBM-bound: both orientations read A with
// coalesced 128-bit loads straight from its one layout -- N: lanes own rows and accumulate over
// columns (no cross-lane reduction); T: a warp owns a column, lanes stride down it and partial
// sums are combined with warp shuffles in a fixed order (deterministic, no atomics).
// f32 products are accumulated in f32 over at most 256 (N) / 64 (T) terms and then carried in
// f64, which keeps 1e-5 relative accuracy at K = 131072.
#include "common.hpp"
#include "cplx.cuh"

namespace jets {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };

struct GemvParams {
  const DBlock* blocks;
  const int32_t* row_ptr;   // [ngroups+1] entries of each output group
  const int32_t* tile_ptr;  // [ngroups+1] CTA tiles of each output group
  int32_t ngroups;
  int32_t acc;              // ACC_SET / ACC_ADD / ACC_SUB
  const char* in;
  char* out;
  // N orientation with few row tiles (a block-ROW shard of a wide operator: 4 x 32 blocks give 64 CTAs for 148 SMs):
  // `ksplit` CTAs share a row tile, each summing a contiguous part of the group's blocks into f64 partials
  // [part][tile][TM]; gemv_n_finish_kernel adds the parts in order (deterministic) and stores.
  int32_t ksplit;
  double* partials;
  int32_t ntiles;
};

__device__ __forceinline__ int find_group(const int32_t* tile_ptr, int ngroups, int tile) {
  int lo = 0, hi = ngroups - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_ptr[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename T>
__device__ __forceinline__ void finish_store(T* o, double v, int acc) {
  if (acc == ACC_SET) *o = (T)v;
  else if (acc == ACC_ADD) *o = (T)((double)*o + v);
  else *o = (T)((double)*o - v);
}

// ---------------------------------------------------------------- N: out = A * in ---------
// CTA = TM rows of one output group (TM = 32*V); the 8 warps split the columns of every block.
template <typename T>
__global__ void __launch_bounds__(kThreads) gemv_n_kernel(const GemvParams P) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  constexpr int TM = 32 * V;
  constexpr int CU = 8;  // columns in flight per lane
  __shared__ double red[kWarps][TM];
  const int gtile = P.ksplit > 1 ? (int)(blockIdx.x / P.ksplit) : (int)blockIdx.x;
  const int part_k = P.ksplit > 1 ? (int)(blockIdx.x % P.ksplit) : 0;
  const int g = find_group(P.tile_ptr, P.ngroups, gtile);
  const int tile = gtile - P.tile_ptr[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int e_begin = P.row_ptr[g], e_end = P.row_ptr[g + 1];
  if (P.ksplit > 1) {        // this CTA's contiguous share of the group's blocks
    const int ne = e_end - e_begin;
    const int a = e_begin + (int)((int64_t)ne * part_k / P.ksplit), b2 = e_begin + (int)((int64_t)ne * (part_k + 1) / P.ksplit);
    e_begin = a; e_end = b2;
  }
  const int64_t i0 = (int64_t)tile * TM + lane * V;  // first row owned by this lane
  double acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.0;
  int64_t out_off = 0;
  int out_len = 0;
  for (int e = e_begin; e < e_end; ++e) {
    const DBlock b = P.blocks[e];
    out_off = b.out_off;
    out_len = b.rows;
    const T* A = reinterpret_cast<const T*>(b.A);
    const T* x = reinterpret_cast<const T*>(P.in) + b.in_off;
    // this warp's column range
    const int cw = (b.cols + kWarps - 1) / kWarps;
    const int j_begin = warp * cw;
    const int j_end = min(b.cols, j_begin + cw);
    // warp-uniform: the whole row tile is inside the matrix and every column is 16B aligned
    const bool vec_ok = ((int64_t)(tile + 1) * TM <= b.rows) && ((b.lda % V) == 0) &&
                        ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    T part[V];
#pragma unroll
    for (int j = 0; j < V; ++j) part[j] = T(0);
    for (int j0 = j_begin; j0 < j_end; j0 += 32) {
      // lane l caches x[j0 + l]; columns are broadcast with shuffles
      const int jl = j0 + lane;
      const T xl = (jl < j_end) ? x[jl] : T(0);
      const int nj = min(32, j_end - j0);
      if (vec_ok) {
        int jj = 0;
        for (; jj + CU <= nj; jj += CU) {
          Vec a[CU];
#pragma unroll
          for (int u = 0; u < CU; ++u)
            a[u] = __ldcs(reinterpret_cast<const Vec*>(A + (int64_t)(j0 + jj + u) * b.lda + i0));
#pragma unroll
          for (int u = 0; u < CU; ++u) {
            const T xv = __shfl_sync(0xffffffffu, xl, jj + u);
            const T* as = reinterpret_cast<const T*>(&a[u]);
#pragma unroll
            for (int k = 0; k < V; ++k) part[k] += as[k] * xv;
          }
        }
        for (; jj < nj; ++jj) {
          const Vec a = __ldcs(reinterpret_cast<const Vec*>(A + (int64_t)(j0 + jj) * b.lda + i0));
          const T xv = __shfl_sync(0xffffffffu, xl, jj);
          const T* as = reinterpret_cast<const T*>(&a);
#pragma unroll
          for (int k = 0; k < V; ++k) part[k] += as[k] * xv;
        }
      } else {
        for (int jj = 0; jj < nj; ++jj) {
          const T xv = __shfl_sync(0xffffffffu, xl, jj);
#pragma unroll
          for (int k = 0; k < V; ++k)
            if (i0 + k < b.rows) part[k] += A[(int64_t)(j0 + jj) * b.lda + i0 + k] * xv;
        }
      }
      if (sizeof(T) == 4 && ((j0 - j_begin) & 255) == 224) {  // fold f32 partial every 256 cols
#pragma unroll
        for (int k = 0; k < V; ++k) { acc[k] += (double)part[k]; part[k] = T(0); }
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += (double)part[k];
  }
#pragma unroll
  for (int k = 0; k < V; ++k) red[warp][lane * V + k] = acc[k];
  __syncthreads();
  // fixed-order cross-warp sum; threads 0..TM-1 each finish one row
  if (threadIdx.x < TM) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
    if (P.ksplit > 1) {
      P.partials[((size_t)part_k * P.ntiles + gtile) * TM + threadIdx.x] = s;
      return;
    }
    const int64_t i = (int64_t)tile * TM + threadIdx.x;
    if (i < out_len) finish_store(reinterpret_cast<T*>(P.out) + out_off + i, s, P.acc);
  }
}

// Second pass of the split N orientation: parts added in order, one thread per output row.
template <typename T>
__global__ void __launch_bounds__(32 * VecOf<T>::V) gemv_n_finish_kernel(const GemvParams P) {
  constexpr int TM = 32 * VecOf<T>::V;
  const int gtile = blockIdx.x;
  const int g = find_group(P.tile_ptr, P.ngroups, gtile);
  const int tile = gtile - P.tile_ptr[g];
  const DBlock b = P.blocks[P.row_ptr[g]];
  double s = 0.0;
  for (int k = 0; k < P.ksplit; ++k) s += P.partials[((size_t)k * P.ntiles + gtile) * TM + threadIdx.x];
  const int64_t i = (int64_t)tile * TM + threadIdx.x;
  if (i < b.rows) finish_store(reinterpret_cast<T*>(P.out) + b.out_off + i, s, P.acc);
}

// ---------------------------------------------------------------- T: out = A' * in --------
// CTA = TN columns of one output group; warp w owns columns {w, w+8, ...} of the tile; the input
// block is staged once per CTA in shared memory.
template <typename T>
__global__ void __launch_bounds__(kThreads) gemv_t_kernel(const GemvParams P) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  constexpr int CW = 4;             // columns per warp
  constexpr int TN = kWarps * CW;   // columns per CTA
  constexpr int U = 8;              // vectors per lane per row chunk
  constexpr int RC = 32 * U * V;    // rows per chunk
  __shared__ __align__(16) T xs[RC];
  const int g = find_group(P.tile_ptr, P.ngroups, blockIdx.x);
  const int tile = blockIdx.x - P.tile_ptr[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e_begin = P.row_ptr[g], e_end = P.row_ptr[g + 1];
  double acc[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) acc[c] = 0.0;
  int64_t out_off = 0;
  int out_len = 0;
  for (int e = e_begin; e < e_end; ++e) {
    const DBlock b = P.blocks[e];
    out_off = b.out_off;
    out_len = b.cols;
    const T* A = reinterpret_cast<const T*>(b.A);
    const T* x = reinterpret_cast<const T*>(P.in) + b.in_off;
    const bool vec_ok = ((b.lda % V) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    for (int r0 = 0; r0 < b.rows; r0 += RC) {
      const int nr = min(RC, b.rows - r0);
      __syncthreads();
      for (int i = threadIdx.x; i < RC; i += kThreads) xs[i] = (i < nr) ? x[r0 + i] : T(0);
      __syncthreads();
      Vec xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = reinterpret_cast<const Vec*>(xs)[u * 32 + lane];
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const int j = tile * TN + c * kWarps + warp;
        if (j >= b.cols) continue;
        const T* col = A + (int64_t)j * b.lda + r0;
        T part = T(0);
        if (vec_ok && nr == RC) {
          Vec a[U];
#pragma unroll
          for (int u = 0; u < U; ++u) a[u] = __ldcs(reinterpret_cast<const Vec*>(col) + u * 32 + lane);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const T* as = reinterpret_cast<const T*>(&a[u]);
            const T* bs = reinterpret_cast<const T*>(&xv[u]);
#pragma unroll
            for (int k = 0; k < V; ++k) part += as[k] * bs[k];
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const T* bs = reinterpret_cast<const T*>(&xv[u]);
#pragma unroll
            for (int k = 0; k < V; ++k) {
              const int i = (u * 32 + lane) * V + k;
              if (i < nr) part += col[i] * bs[k];
            }
          }
        }
        acc[c] += (double)part;
      }
    }
  }
  // one fixed-order shuffle tree per owned column
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    double v = acc[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int j = tile * TN + c * kWarps + warp;
    if (lane == 0 && j < out_len) finish_store(reinterpret_cast<T*>(P.out) + out_off + j, v, P.acc);
  }
}

// ---------------------------------------------------------------- complex eltypes ---------
// d = A m and m = A' d with A' the CONJUGATE transpose (_matmul_df'! src/Jets.jl:574: mul!(m, A', d)) on
// interleaved ComplexF32 / ComplexF64 storage.  Same decomposition as the real kernels: N -- a lane owns V
// consecutive rows (one 128-bit load per column) and the 8 warps split the columns; T -- a warp owns a column and
// the lanes stride down it.  Products follow Julia's complex multiply (cplx.cuh); every partial is carried in f64.
template <typename R> struct CVecOf;
template <> struct CVecOf<float>  { using type = float4;  static constexpr int V = 2; };
template <> struct CVecOf<double> { using type = double2; static constexpr int V = 1; };

template <typename R>
__device__ __forceinline__ void cfinish_store(Cx<R>* o, double re, double im, int acc) {
  if (acc == ACC_SET) { o->re = (R)re; o->im = (R)im; }
  else if (acc == ACC_ADD) { o->re = (R)((double)o->re + re); o->im = (R)((double)o->im + im); }
  else { o->re = (R)((double)o->re - re); o->im = (R)((double)o->im - im); }
}

template <typename R>
__global__ void __launch_bounds__(kThreads) cgemv_n_kernel(const GemvParams P) {
  using Vec = typename CVecOf<R>::type;
  using Z = Cx<R>;
  constexpr int V = CVecOf<R>::V;
  constexpr int TM = 32 * V;
  constexpr int CU = 4;
  __shared__ double red[kWarps][TM][2];
  const int g = find_group(P.tile_ptr, P.ngroups, (int)blockIdx.x);
  const int tile = (int)blockIdx.x - P.tile_ptr[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e_begin = P.row_ptr[g], e_end = P.row_ptr[g + 1];
  const int64_t i0 = (int64_t)tile * TM + lane * V;
  double are[V], aim[V];
#pragma unroll
  for (int k = 0; k < V; ++k) are[k] = aim[k] = 0.0;
  int64_t out_off = 0;
  int out_len = 0;
  for (int e = e_begin; e < e_end; ++e) {
    const DBlock b = P.blocks[e];
    out_off = b.out_off;
    out_len = b.rows;
    const Z* A = reinterpret_cast<const Z*>(b.A);
    const Z* x = reinterpret_cast<const Z*>(P.in) + b.in_off;
    const int cw = (b.cols + kWarps - 1) / kWarps;
    const int j_begin = warp * cw;
    const int j_end = min(b.cols, j_begin + cw);
    const bool vec_ok = ((int64_t)(tile + 1) * TM <= b.rows) && ((b.lda % V) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    for (int j0 = j_begin; j0 < j_end; j0 += 32) {
      const int jl = j0 + lane;
      const Z xl = (jl < j_end) ? x[jl] : Z(R(0));
      const int nj = min(32, j_end - j0);
      Z part[V];
#pragma unroll
      for (int k = 0; k < V; ++k) part[k] = Z(R(0));
      if (vec_ok) {
        int jj = 0;
        for (; jj + CU <= nj; jj += CU) {
          Vec a[CU];
#pragma unroll
          for (int u = 0; u < CU; ++u) a[u] = __ldcs(reinterpret_cast<const Vec*>(A + (int64_t)(j0 + jj + u) * b.lda + i0));
#pragma unroll
          for (int u = 0; u < CU; ++u) {
            const Z xv(__shfl_sync(0xffffffffu, xl.re, jj + u), __shfl_sync(0xffffffffu, xl.im, jj + u));
            const Z* as = reinterpret_cast<const Z*>(&a[u]);
#pragma unroll
            for (int k = 0; k < V; ++k) part[k] = part[k] + as[k] * xv;
          }
        }
        for (; jj < nj; ++jj) {
          const Vec a = __ldcs(reinterpret_cast<const Vec*>(A + (int64_t)(j0 + jj) * b.lda + i0));
          const Z xv(__shfl_sync(0xffffffffu, xl.re, jj), __shfl_sync(0xffffffffu, xl.im, jj));
          const Z* as = reinterpret_cast<const Z*>(&a);
#pragma unroll
          for (int k = 0; k < V; ++k) part[k] = part[k] + as[k] * xv;
        }
      } else {
        for (int jj = 0; jj < nj; ++jj) {
          const Z xv(__shfl_sync(0xffffffffu, xl.re, jj), __shfl_sync(0xffffffffu, xl.im, jj));
#pragma unroll
          for (int k = 0; k < V; ++k)
            if (i0 + k < b.rows) part[k] = part[k] + A[(int64_t)(j0 + jj) * b.lda + i0 + k] * xv;
        }
      }
#pragma unroll
      for (int k = 0; k < V; ++k) { are[k] += (double)part[k].re; aim[k] += (double)part[k].im; }   // folded every 32 columns
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) { red[warp][lane * V + k][0] = are[k]; red[warp][lane * V + k][1] = aim[k]; }
  __syncthreads();
  if (threadIdx.x < TM) {
    double sr = 0.0, si = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { sr += red[w][threadIdx.x][0]; si += red[w][threadIdx.x][1]; }
    const int64_t i = (int64_t)tile * TM + threadIdx.x;
    if (i < out_len) cfinish_store(reinterpret_cast<Z*>(P.out) + out_off + i, sr, si, P.acc);
  }
}

template <typename R>
__global__ void __launch_bounds__(kThreads) cgemv_t_kernel(const GemvParams P) {
  using Vec = typename CVecOf<R>::type;
  using Z = Cx<R>;
  constexpr int V = CVecOf<R>::V;
  constexpr int CW = 4;
  constexpr int TN = kWarps * CW;
  constexpr int U = 4;
  constexpr int RC = 32 * U * V;    // rows per chunk
  __shared__ __align__(16) Z xs[RC];
  const int g = find_group(P.tile_ptr, P.ngroups, (int)blockIdx.x);
  const int tile = (int)blockIdx.x - P.tile_ptr[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e_begin = P.row_ptr[g], e_end = P.row_ptr[g + 1];
  double are[CW], aim[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) are[c] = aim[c] = 0.0;
  int64_t out_off = 0;
  int out_len = 0;
  for (int e = e_begin; e < e_end; ++e) {
    const DBlock b = P.blocks[e];
    out_off = b.out_off;
    out_len = b.cols;
    const Z* A = reinterpret_cast<const Z*>(b.A);
    const Z* x = reinterpret_cast<const Z*>(P.in) + b.in_off;
    const bool vec_ok = ((b.lda % V) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    for (int r0 = 0; r0 < b.rows; r0 += RC) {
      const int nr = min(RC, b.rows - r0);
      __syncthreads();
      for (int i = threadIdx.x; i < RC; i += kThreads) xs[i] = (i < nr) ? x[r0 + i] : Z(R(0));
      __syncthreads();
      Vec xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = reinterpret_cast<const Vec*>(xs)[u * 32 + lane];
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const int j = tile * TN + c * kWarps + warp;
        if (j >= b.cols) continue;
        const Z* col = A + (int64_t)j * b.lda + r0;
        Z part(R(0));
        if (vec_ok && nr == RC) {
          Vec a[U];
#pragma unroll
          for (int u = 0; u < U; ++u) a[u] = __ldcs(reinterpret_cast<const Vec*>(col) + u * 32 + lane);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const Z* as = reinterpret_cast<const Z*>(&a[u]);
            const Z* bs = reinterpret_cast<const Z*>(&xv[u]);
#pragma unroll
            for (int k = 0; k < V; ++k) part = part + conj(as[k]) * bs[k];
          }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const Z* bs = reinterpret_cast<const Z*>(&xv[u]);
#pragma unroll
            for (int k = 0; k < V; ++k) {
              const int i = (u * 32 + lane) * V + k;
              if (i < nr) part = part + conj(col[i]) * bs[k];
            }
          }
        }
        are[c] += (double)part.re;
        aim[c] += (double)part.im;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CW; ++c) {
    double vr = are[c], vi = aim[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { vr += __shfl_xor_sync(0xffffffffu, vr, o); vi += __shfl_xor_sync(0xffffffffu, vi, o); }
    const int j = tile * TN + c * kWarps + warp;
    if (lane == 0 && j < out_len) cfinish_store(reinterpret_cast<Z*>(P.out) + out_off + j, vr, vi, P.acc);
  }
}

}  // namespace

void gemv_tile_count(int dtype, bool trans, int32_t out_len, int32_t* ntiles) {
  const int V = dtype == JETS_F32 ? 4 : dtype == JETS_C128 ? 1 : 2;
  const int per = trans ? (kWarps * 4) : 32 * V;
  *ntiles = (out_len + per - 1) / per;
}

void launch_gemv(const Step& st, int dtype, const char* in, char* out, cudaStream_t s) {
  if (st.n_out_rows == 0) return;
  GemvParams P;
  P.blocks = st.d_dblocks;
  P.row_ptr = st.d_row_ptr;
  P.tile_ptr = st.d_row_ptr + (st.n_out_rows + 1);
  P.ngroups = st.n_out_rows;
  P.acc = st.acc;
  P.in = in;
  P.out = out;
  const bool trans = st.dblocks[0].trans != 0;
  P.ksplit = (!trans && st.gemv_ksplit > 1 && st.gemv_partials) ? st.gemv_ksplit : 1;
  P.partials = st.gemv_partials;
  P.ntiles = (int32_t)st.gemv_tiles;
  const unsigned grid = (unsigned)(st.gemv_tiles * P.ksplit);
  if (is_cplx(dtype)) {       // never split (emit_gemv leaves gemv_ksplit at 1 for complex eltypes)
    if (dtype == JETS_C64) {
      if (trans) cgemv_t_kernel<float><<<grid, kThreads, 0, s>>>(P);
      else cgemv_n_kernel<float><<<grid, kThreads, 0, s>>>(P);
    } else {
      if (trans) cgemv_t_kernel<double><<<grid, kThreads, 0, s>>>(P);
      else cgemv_n_kernel<double><<<grid, kThreads, 0, s>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    count_launch();
    return;
  }
  if (dtype == JETS_F32) {
    if (trans) gemv_t_kernel<float><<<grid, kThreads, 0, s>>>(P);
    else gemv_n_kernel<float><<<grid, kThreads, 0, s>>>(P);
  } else {
    if (trans) gemv_t_kernel<double><<<grid, kThreads, 0, s>>>(P);
    else gemv_n_kernel<double><<<grid, kThreads, 0, s>>>(P);
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  if (P.ksplit > 1) {
    if (dtype == JETS_F32) gemv_n_finish_kernel<float><<<(unsigned)st.gemv_tiles, 128, 0, s>>>(P);
    else gemv_n_finish_kernel<double><<<(unsigned)st.gemv_tiles, 64, 0, s>>>(P);
    CUDA_TRY(cudaGetLastError());
    count_launch();
  }
}

}  // namespace jets
