// BlockArray vector-space kernels: fill!/rand, broadcast updates (src/Jets.jl:880-911) and the
// dot/norm/extrema reductions (src/Jets.jl:834-878).  Storage is one flat device buffer, so a
// BlockArray broadcast is a single streaming pass; reductions are two-pass, fixed-order, f64
// accumulated, warp-shuffle trees -- no atomics, bit-reproducible run to run.
#include <algorithm>
#include <cstdlib>
#include "common.hpp"
#include "cplx.cuh"

namespace jets {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };

inline unsigned grid_for(int64_t nvec, int per_thread) {
  int64_t g = (nvec + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  const int64_t cap = (int64_t)ctx().sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ------------------------------------------------------------------ fill -----------------
template <typename T>
__global__ void __launch_bounds__(kThreads) fill_kernel(T* __restrict__ p, int64_t n, T a) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  // head: unaligned prefix
  const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
  int64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
  if (head > n) head = n;
  const int64_t nvec = (n - head) / V;
  Vec v;
  T* vs = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int j = 0; j < V; ++j) vs[j] = a;
  Vec* pv = reinterpret_cast<Vec*>(p + head);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) pv[i] = v;
  const int64_t tail0 = head + nvec * V;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid < head) p[gid] = a;
  if (gid < n - tail0) p[tail0 + gid] = a;
}

// ------------------------------------------------------------------ Philox4x32-10 --------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// One Philox block per element index: value depends only on (seed, logical index, dist), so any
// block / multi-GPU partition of the same logical vector draws the same numbers.
template <typename T>
__global__ void __launch_bounds__(kThreads) rand_kernel(T* __restrict__ p, int64_t n, uint64_t seed,
                                                        uint64_t off, int dist) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t idx = off + (uint64_t)i;
    uint32_t r[4];
    philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)dist, 0u, (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    const double u1 = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) * (1.0 / 9007199254740992.0);
    if (dist == 0) {
      if (sizeof(T) == 4) p[i] = (T)((float)(r[0] >> 8) * (1.0f / 16777216.0f));
      else p[i] = (T)u1;
    } else {
      const double u2 = ((double)(r[2] >> 5) * 67108864.0 + (double)(r[3] >> 6)) * (1.0 / 9007199254740992.0);
      const double z = sqrt(-2.0 * log(1.0 - u1)) * cospi(2.0 * u2);
      p[i] = (T)z;
    }
  }
}

// ------------------------------------------------------------------ lincomb / hadamard ----
struct LinArgs {
  const void* x[4];
  double c[4];
};

template <typename T, int K, bool VEC>
__global__ void __launch_bounds__(kThreads) lincomb_kernel(T* __restrict__ out, int64_t n, LinArgs a) {
  pdl_enter();
  using Vec = typename VecOf<T>::type;
  constexpr int V = VEC ? VecOf<T>::V : 1;
  const int64_t nvec = VEC ? n / V : n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  T c[K];
#pragma unroll
  for (int k = 0; k < K; ++k) c[k] = (T)a.c[k];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    T r[V];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      T xv[V];
      if (VEC) {
        const Vec v = reinterpret_cast<const Vec*>(a.x[k])[i];
        const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int j = 0; j < V; ++j) xv[j] = vs[j];
      } else {
        xv[0] = reinterpret_cast<const T*>(a.x[k])[i];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) r[j] = (k == 0) ? c[0] * xv[j] : r[j] + c[k] * xv[j];
    }
    if (VEC) {
      Vec v;
      T* vs = reinterpret_cast<T*>(&v);
#pragma unroll
      for (int j = 0; j < V; ++j) vs[j] = r[j];
      reinterpret_cast<Vec*>(out)[i] = v;
    } else {
      out[i] = r[0];
    }
  }
  if (VEC) {  // scalar tail
    const int64_t t0 = nvec * V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - t0) {
      T r = T(0);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const T xv = reinterpret_cast<const T*>(a.x[k])[t0 + gid];
        r = (k == 0) ? c[0] * xv : r + c[k] * xv;
      }
      out[t0 + gid] = r;
    }
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) hadamard_kernel(T* __restrict__ out, const T* x, const T* y,
                                                            int64_t n) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VEC ? VecOf<T>::V : 1;
  const int64_t nvec = VEC ? n / V : n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    if (VEC) {
      const Vec a = reinterpret_cast<const Vec*>(x)[i];
      const Vec b = reinterpret_cast<const Vec*>(y)[i];
      Vec r;
      const T* as = reinterpret_cast<const T*>(&a);
      const T* bs = reinterpret_cast<const T*>(&b);
      T* rs = reinterpret_cast<T*>(&r);
#pragma unroll
      for (int j = 0; j < V; ++j) rs[j] = as[j] * bs[j];
      reinterpret_cast<Vec*>(out)[i] = r;
    } else {
      out[i] = x[i] * y[i];
    }
  }
  if (VEC) {
    const int64_t t0 = nvec * V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - t0) out[t0 + gid] = x[t0 + gid] * y[t0 + gid];
  }
}

// out = A*x + B*y with A,B read from device scalars (or constants); y may be null.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) axpby_dev_kernel(T* __restrict__ out, int64_t n,
                                                             const double* sa, double ca, int af,
                                                             const T* x, const double* sb, double cb,
                                                             int bf, const T* y) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VEC ? VecOf<T>::V : 1;
  pdl_enter();   // the device scalars are produced by the kernels before this one
  double A = sa ? *sa : ca;
  if (af & JETS_COEF_INV) A = 1.0 / A;
  if (af & JETS_COEF_NEG) A = -A;
  double B = sb ? *sb : cb;
  if (bf & JETS_COEF_INV) B = 1.0 / B;
  if (bf & JETS_COEF_NEG) B = -B;
  const T a = (T)A, b = (T)B;
  const int64_t nvec = VEC ? n / V : n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    if (VEC) {
      const Vec xv = reinterpret_cast<const Vec*>(x)[i];
      const T* xs = reinterpret_cast<const T*>(&xv);
      Vec r;
      T* rs = reinterpret_cast<T*>(&r);
      if (y) {
        const Vec yv = reinterpret_cast<const Vec*>(y)[i];
        const T* ys = reinterpret_cast<const T*>(&yv);
#pragma unroll
        for (int j = 0; j < V; ++j) rs[j] = a * xs[j] + b * ys[j];
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) rs[j] = a * xs[j];
      }
      reinterpret_cast<Vec*>(out)[i] = r;
    } else {
      out[i] = y ? a * x[i] + b * y[i] : a * x[i];
    }
  }
  if (VEC) {
    const int64_t t0 = nvec * V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - t0) out[t0 + gid] = y ? a * x[t0 + gid] + b * y[t0 + gid] : a * x[t0 + gid];
  }
}

// Two such updates in ONE pass (CG: x += a p, r -= a q; LSQR: x += t1 w, w = v/alpha - t2 w): every input element is
// read before either output element is written, so an output may alias any input.  Same arithmetic per update as
// axpby_dev_kernel (one rounding per operation), so the results are bit-identical to two separate calls.
struct AxpbyPair {
  void* out[2]; const void* x[2]; const void* y[2];
  const double* sa[2]; const double* sb[2];
  double ca[2], cb[2];
  int af[2], bf[2];
};
__device__ __forceinline__ double coef_of(const double* s, double c, int f) {
  double v = s ? *s : c;
  if (f & JETS_COEF_INV) v = 1.0 / v;
  if (f & JETS_COEF_NEG) v = -v;
  return v;
}
template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) axpby_pair_kernel(const AxpbyPair P, int64_t n) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VEC ? VecOf<T>::V : 1;
  pdl_enter();
  const T a0 = (T)coef_of(P.sa[0], P.ca[0], P.af[0]), b0 = (T)coef_of(P.sb[0], P.cb[0], P.bf[0]);
  const T a1 = (T)coef_of(P.sa[1], P.ca[1], P.af[1]), b1 = (T)coef_of(P.sb[1], P.cb[1], P.bf[1]);
  const int64_t nvec = VEC ? n / V : n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    if (VEC) {
      const Vec x0 = reinterpret_cast<const Vec*>(P.x[0])[i], y0 = reinterpret_cast<const Vec*>(P.y[0])[i];
      const Vec x1 = reinterpret_cast<const Vec*>(P.x[1])[i], y1 = reinterpret_cast<const Vec*>(P.y[1])[i];
      Vec r0, r1;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        reinterpret_cast<T*>(&r0)[j] = a0 * reinterpret_cast<const T*>(&x0)[j] + b0 * reinterpret_cast<const T*>(&y0)[j];
        reinterpret_cast<T*>(&r1)[j] = a1 * reinterpret_cast<const T*>(&x1)[j] + b1 * reinterpret_cast<const T*>(&y1)[j];
      }
      reinterpret_cast<Vec*>(P.out[0])[i] = r0;
      reinterpret_cast<Vec*>(P.out[1])[i] = r1;
    } else {
      const T x0 = reinterpret_cast<const T*>(P.x[0])[i], y0 = reinterpret_cast<const T*>(P.y[0])[i];
      const T x1 = reinterpret_cast<const T*>(P.x[1])[i], y1 = reinterpret_cast<const T*>(P.y[1])[i];
      reinterpret_cast<T*>(P.out[0])[i] = a0 * x0 + b0 * y0;
      reinterpret_cast<T*>(P.out[1])[i] = a1 * x1 + b1 * y1;
    }
  }
  if (VEC) {
    const int64_t t0 = nvec * V;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n - t0) {
      const int64_t i = t0 + gid;
      const T x0 = reinterpret_cast<const T*>(P.x[0])[i], y0 = reinterpret_cast<const T*>(P.y[0])[i];
      const T x1 = reinterpret_cast<const T*>(P.x[1])[i], y1 = reinterpret_cast<const T*>(P.y[1])[i];
      reinterpret_cast<T*>(P.out[0])[i] = a0 * x0 + b0 * y0;
      reinterpret_cast<T*>(P.out[1])[i] = a1 * x1 + b1 * y1;
    }
  }
}

// ------------------------------------------------------------------ reductions -----------
enum RKind : int { R_DOT = 0, R_SUMSQ, R_SUMABS, R_NNZ, R_MAXABS, R_MINABS, R_SUMPOW, R_MIN, R_MAX };

template <int KIND>
__device__ __forceinline__ double r_identity() {
  if (KIND == R_MAXABS) return 0.0;
  if (KIND == R_MINABS || KIND == R_MIN) return __longlong_as_double(0x7ff0000000000000LL);
  if (KIND == R_MAX) return __longlong_as_double(0xfff0000000000000LL);
  return 0.0;
}
template <int KIND>
__device__ __forceinline__ double r_combine(double a, double b) {
  if (KIND == R_MAXABS || KIND == R_MAX) return fmax(a, b);
  if (KIND == R_MINABS || KIND == R_MIN) return fmin(a, b);
  return a + b;
}
template <int KIND>
__device__ __forceinline__ double r_map(double x, double y, double p) {
  switch (KIND) {
    case R_DOT: return x * y;
    case R_SUMSQ: return x * x;
    case R_SUMABS: return fabs(x);
    case R_NNZ: return x != 0.0 ? 1.0 : 0.0;
    case R_MAXABS: case R_MINABS: return fabs(x);
    case R_SUMPOW: return pow(fabs(x), p);
    default: return x;
  }
}
template <int KIND>
__device__ __forceinline__ double block_reduce(double v) {
  __shared__ double sh[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = r_combine<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = r_identity<KIND>();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) r = r_combine<KIND>(r, sh[i]);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// ONE launch: CTA b reduces the contiguous chunk [b*chunk, (b+1)*chunk) -> partial[b]; the CTA that draws the last
// ticket then combines the partials in index order (thread t takes partial[t], partial[t + 256], ...: the order does not
// depend on which CTA finishes last, so the result is deterministic) and applies `finish`: 0 none, 1 sqrt, 2 ^(1/p).
// The ticket wraps back to 0 by itself (atomicInc).  Reductions share the context's partial buffer: one at a time.
template <typename T, int KIND>
__global__ void __launch_bounds__(kThreads) reduce_pass1(const T* __restrict__ x, const T* __restrict__ y,
                                                         int64_t n, int64_t chunk, double p,
                                                         double* partial, unsigned int* ticket, int finish,
                                                         double* __restrict__ out) {
  using Vec = typename VecOf<T>::type;
  constexpr int V = VecOf<T>::V;
  pdl_enter();
  const int64_t b0 = (int64_t)blockIdx.x * chunk;
  int64_t b1 = b0 + chunk;
  if (b1 > n) b1 = n;
  double acc[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) acc[u] = r_identity<KIND>();
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x + b0) & 15) == 0) &&
                      (KIND != R_DOT || (reinterpret_cast<uintptr_t>(y + b0) & 15) == 0);
  int64_t done = b0;
  if (vec_ok) {
    const int64_t nvec = (b1 - b0) / V;
    const Vec* xv = reinterpret_cast<const Vec*>(x + b0);
    const Vec* yv = reinterpret_cast<const Vec*>(KIND == R_DOT ? y + b0 : x + b0);
    int64_t i = threadIdx.x;
    for (; i + (kUnroll - 1) * kThreads < nvec; i += kUnroll * kThreads) {
      Vec a[kUnroll], b[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        a[u] = xv[i + u * kThreads];
        if (KIND == R_DOT) b[u] = yv[i + u * kThreads];
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const T* as = reinterpret_cast<const T*>(&a[u]);
        const T* bs = reinterpret_cast<const T*>(&b[u]);
#pragma unroll
        for (int j = 0; j < V; ++j)
          acc[u] = r_combine<KIND>(acc[u], r_map<KIND>((double)as[j], KIND == R_DOT ? (double)bs[j] : 0.0, p));
      }
    }
    for (; i < nvec; i += kThreads) {
      const Vec a = xv[i];
      Vec b = a;
      if (KIND == R_DOT) b = yv[i];
      const T* as = reinterpret_cast<const T*>(&a);
      const T* bs = reinterpret_cast<const T*>(&b);
#pragma unroll
      for (int j = 0; j < V; ++j)
        acc[0] = r_combine<KIND>(acc[0], r_map<KIND>((double)as[j], (double)bs[j], p));
    }
    done = b0 + nvec * V;
  }
  for (int64_t i = done + threadIdx.x; i < b1; i += kThreads)
    acc[0] = r_combine<KIND>(acc[0], r_map<KIND>((double)x[i], KIND == R_DOT ? (double)y[i] : 0.0, p));
  double v = acc[0];
#pragma unroll
  for (int u = 1; u < kUnroll; ++u) v = r_combine<KIND>(v, acc[u]);
  v = block_reduce<KIND>(v);
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = v;
    __threadfence();                                   // the partial is visible before the ticket is drawn
    is_last = ticket ? atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1 : 0;   // no ticket: a second launch finishes (A/B)
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int np = (int)gridDim.x;
  double r = r_identity<KIND>();
  for (int i = threadIdx.x; i < np; i += kThreads) r = r_combine<KIND>(r, __ldcg(partial + i));
  r = block_reduce<KIND>(r);
  if (threadIdx.x == 0) {
    if (finish == 1) r = sqrt(r);
    else if (finish == 2) r = pow(r, 1.0 / p);
    out[0] = r;
  }
}

// the finish as a launch of its own (JETS_B200_REDUCE_TWO_PASS=1: the A/B baseline of the single-launch reduction)
template <int KIND>
__global__ void __launch_bounds__(kThreads) reduce_pass2(const double* __restrict__ partial, int np,
                                                         int finish, double p, double* __restrict__ out) {
  double v = r_identity<KIND>();
  for (int i = threadIdx.x; i < np; i += kThreads) v = r_combine<KIND>(v, partial[i]);
  v = block_reduce<KIND>(v);
  if (threadIdx.x == 0) {
    if (finish == 1) v = sqrt(v);
    else if (finish == 2) v = pow(v, 1.0 / p);
    out[0] = v;
  }
}

template <typename T, int KIND>
void reduce_launch(const void* x, const void* y, int64_t n, double p, int finish, double* dev_out,
                   cudaStream_t s) {
  Context& c = ctx();
  constexpr int V = VecOf<T>::V;
  // chunk: multiple of V*kThreads*kUnroll elements, <= 8 CTAs per SM
  const int64_t quantum = (int64_t)V * kThreads * kUnroll;
  int64_t nb = (n + quantum - 1) / quantum;
  const int64_t cap = (int64_t)c.sm_count * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  int64_t chunk = ((n + nb - 1) / nb + quantum - 1) / quantum * quantum;
  if (chunk < quantum) chunk = quantum;
  nb = n > 0 ? (n + chunk - 1) / chunk : 1;
  JETS_CHECK((size_t)nb <= c.dev_scratch_elems, JETS_ERR_INVALID, "reduction scratch too small");
  static const bool two_pass = getenv("JETS_B200_REDUCE_TWO_PASS") && atoi(getenv("JETS_B200_REDUCE_TWO_PASS"));
  unsigned int* ticket = two_pass ? nullptr : reinterpret_cast<unsigned int*>(c.dev_scratch + c.dev_scratch_elems + 60);
  launch_pdl(reduce_pass1<T, KIND>, (unsigned)nb, kThreads, s, dev_out, sizeof(double), (const T*)x, (const T*)y, n, chunk, p,
             c.dev_scratch, ticket, finish, dev_out);
  if (two_pass) reduce_pass2<KIND><<<1, kThreads, 0, s>>>(c.dev_scratch, (int)nb, finish, p, dev_out);
  CUDA_TRY(cudaGetLastError());
  count_launch(two_pass ? 2 : 1);
}

template <typename T>
void reduce_dispatch(int kind, const void* x, const void* y, int64_t n, double p, double* out,
                     cudaStream_t s) {
  switch (kind) {
    case 0: reduce_launch<T, R_DOT>(x, y, n, p, 0, out, s); break;
    case 1: reduce_launch<T, R_SUMSQ>(x, y, n, p, 1, out, s); break;   // 2-norm
    case 2: reduce_launch<T, R_SUMABS>(x, y, n, p, 0, out, s); break;
    case 3: reduce_launch<T, R_NNZ>(x, y, n, p, 0, out, s); break;
    case 4: reduce_launch<T, R_MAXABS>(x, y, n, p, 0, out, s); break;
    case 5: reduce_launch<T, R_MINABS>(x, y, n, p, 0, out, s); break;
    case 6: reduce_launch<T, R_SUMPOW>(x, y, n, p, 2, out, s); break;
    case 7: reduce_launch<T, R_MIN>(x, y, n, p, 0, out, s); break;
    case 8: reduce_launch<T, R_MAX>(x, y, n, p, 0, out, s); break;
    default: JETS_FAIL(JETS_ERR_INVALID, "bad reduction kind %d", kind);
  }
}

__device__ __forceinline__ double scalar_eval(char op, double x, double y) {
  switch (op) {
    case '+': return x + y;
    case '-': return x - y;
    case '*': return x * y;
    case '/': return x / y;
    case 'n': return -x;
    case 's': return sqrt(x);
    case 'h': return sqrt(x * x + y * y);
    default: return x;
  }
}
// A short straight-line program of scalar operations in ONE launch (the scalar recurrences of a
// CG/LSQR iteration); operations see the results of the ones before them.
__global__ void __launch_bounds__(32) scalar_prog_kernel(const ScalarProg p) {
  __shared__ double av[kMaxScalarProg], bv[kMaxScalarProg], val[kMaxScalarProg];
  pdl_enter();
  const int t = threadIdx.x;
  if (t < kMaxScalarProg) {                 // lanes 0..15 fetch the first operands, lanes 16..31 the second ones: one round trip
    if (t < p.n) av[t] = (p.asrc[t] < 0 && p.a[t]) ? *p.a[t] : 0.0;
  } else {
    const int k = t - kMaxScalarProg;
    if (k < p.n) bv[k] = (p.bsrc[k] < 0 && p.b[k]) ? *p.b[k] : 0.0;
  }
  __syncwarp();
  if (t == 0) {
    for (int i = 0; i < p.n; ++i) {
      const double x = p.asrc[i] < 0 ? av[i] : val[p.asrc[i]];
      const double y = p.bsrc[i] < 0 ? bv[i] : val[p.bsrc[i]];
      val[i] = scalar_eval(p.op[i], x, y);
    }
  }
  __syncwarp();
  if (t < p.n && p.store[t]) *p.out[t] = val[t];
}
__global__ void scalar_op_kernel(double* out, char op, const double* a, const double* b) {
  pdl_enter();
  const double x = a ? *a : 0.0, y = b ? *b : 0.0;
  double r;
  switch (op) {
    case '+': r = x + y; break;
    case '-': r = x - y; break;
    case '*': r = x * y; break;
    case '/': r = x / y; break;
    case 'n': r = -x; break;
    case 's': r = sqrt(x); break;
    case 'h': r = sqrt(x * x + y * y); break;
    default: r = x; break;
  }
  *out = r;
}

}  // namespace

void vec_fill(int dtype, void* p, int64_t n, double a, cudaStream_t s) {
  if (n <= 0) return;
  if (dtype == JETS_F32) fill_kernel<float><<<grid_for(n / 4 + 1, 4), kThreads, 0, s>>>((float*)p, n, (float)a);
  else fill_kernel<double><<<grid_for(n / 2 + 1, 4), kThreads, 0, s>>>((double*)p, n, a);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void vec_rand(int dtype, void* p, int64_t n, uint64_t seed, uint64_t off, int dist, cudaStream_t s) {
  if (n <= 0) return;
  if (dtype == JETS_F32) rand_kernel<float><<<grid_for(n, 8), kThreads, 0, s>>>((float*)p, n, seed, off, dist);
  else rand_kernel<double><<<grid_for(n, 8), kThreads, 0, s>>>((double*)p, n, seed, off, dist);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

template <typename T, bool VEC>
static void lincomb_k(T* out, int64_t n, int k, const LinArgs& a, cudaStream_t s) {
  const unsigned g = grid_for(VEC ? n / VecOf<T>::V + 1 : n, 2);
  switch (k) {
    case 1: launch_pdl(lincomb_kernel<T, 1, VEC>, g, kThreads, s, out, (size_t)n * sizeof(T), out, n, a); break;
    case 2: launch_pdl(lincomb_kernel<T, 2, VEC>, g, kThreads, s, out, (size_t)n * sizeof(T), out, n, a); break;
    case 3: launch_pdl(lincomb_kernel<T, 3, VEC>, g, kThreads, s, out, (size_t)n * sizeof(T), out, n, a); break;
    default: launch_pdl(lincomb_kernel<T, 4, VEC>, g, kThreads, s, out, (size_t)n * sizeof(T), out, n, a); break;
  }
}

void vec_lincomb(int dtype, void* out, int64_t n, int k, const double* c, const void* const* x,
                 cudaStream_t s) {
  if (n <= 0) return;
  JETS_CHECK(k >= 1 && k <= 4, JETS_ERR_INVALID, "lincomb supports 1..4 terms, got %d", k);
  LinArgs a{};
  bool al = aligned16(out);
  for (int i = 0; i < k; ++i) {
    a.x[i] = x[i];
    a.c[i] = c[i];
    al = al && aligned16(x[i]);
  }
  if (dtype == JETS_F32) {
    if (al) lincomb_k<float, true>((float*)out, n, k, a, s);
    else lincomb_k<float, false>((float*)out, n, k, a, s);
  } else {
    if (al) lincomb_k<double, true>((double*)out, n, k, a, s);
    else lincomb_k<double, false>((double*)out, n, k, a, s);
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void vec_hadamard(int dtype, void* out, const void* x, const void* y, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  const bool al = aligned16(out) && aligned16(x) && aligned16(y);
  if (dtype == JETS_F32) {
    const unsigned g = grid_for(n / 4 + 1, 2);
    if (al) hadamard_kernel<float, true><<<g, kThreads, 0, s>>>((float*)out, (const float*)x, (const float*)y, n);
    else hadamard_kernel<float, false><<<grid_for(n, 2), kThreads, 0, s>>>((float*)out, (const float*)x, (const float*)y, n);
  } else {
    const unsigned g = grid_for(n / 2 + 1, 2);
    if (al) hadamard_kernel<double, true><<<g, kThreads, 0, s>>>((double*)out, (const double*)x, (const double*)y, n);
    else hadamard_kernel<double, false><<<grid_for(n, 2), kThreads, 0, s>>>((double*)out, (const double*)x, (const double*)y, n);
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void vec_reduce(int dtype, int kind, const void* x, const void* y, int64_t n, double p,
                double* dev_out, cudaStream_t s) {
  if (dtype == JETS_F32) reduce_dispatch<float>(kind, x, y, n, p, dev_out, s);
  else reduce_dispatch<double>(kind, x, y, n, p, dev_out, s);
}

void scalar_prog(const ScalarProg& prog, cudaStream_t s) {
  if (prog.n <= 0) return;
  ScalarProg p = prog;
  for (int i = 0; i < p.n; ++i) {
    p.asrc[i] = p.bsrc[i] = -1;
    p.store[i] = 1;
    for (int j = 0; j < i; ++j) {           // the latest earlier step that writes the operand
      if (p.a[i] && p.out[j] == p.a[i]) p.asrc[i] = (int8_t)j;
      if (p.b[i] && p.out[j] == p.b[i]) p.bsrc[i] = (int8_t)j;
    }
    for (int j = i + 1; j < p.n; ++j)
      if (p.out[j] == p.out[i]) p.store[i] = 0;     // overwritten later in the same program
  }
  launch_pdl(scalar_prog_kernel, 1u, 32u, s, nullptr, 0, p);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}
void scalar_op(double* out, char op, const double* a, const double* b, cudaStream_t s) {
  launch_pdl(scalar_op_kernel, 1u, 1u, s, out, sizeof(double), out, op, a, b);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void vec_axpby_dev(int dtype, void* out, int64_t n, const double* sa, double ca, int af,
                   const void* x, const double* sb, double cb, int bf, const void* y,
                   cudaStream_t s) {
  if (n <= 0) return;
  const bool al = aligned16(out) && aligned16(x) && (!y || aligned16(y));
  if (dtype == JETS_F32) {
    if (al) launch_pdl(axpby_dev_kernel<float, true>, grid_for(n / 4 + 1, 2), kThreads, s, out, (size_t)n * sizeof(float), (float*)out, n, sa, ca, af, (const float*)x, sb, cb, bf, (const float*)y);
    else launch_pdl(axpby_dev_kernel<float, false>, grid_for(n, 2), kThreads, s, out, (size_t)n * sizeof(float), (float*)out, n, sa, ca, af, (const float*)x, sb, cb, bf, (const float*)y);
  } else {
    if (al) launch_pdl(axpby_dev_kernel<double, true>, grid_for(n / 2 + 1, 2), kThreads, s, out, (size_t)n * sizeof(double), (double*)out, n, sa, ca, af, (const double*)x, sb, cb, bf, (const double*)y);
    else launch_pdl(axpby_dev_kernel<double, false>, grid_for(n, 2), kThreads, s, out, (size_t)n * sizeof(double), (double*)out, n, sa, ca, af, (const double*)x, sb, cb, bf, (const double*)y);
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void vec_axpby_pair_dev(int dtype, int64_t n, void* const out[2], const double* const sa[2], const double ca[2], const int af[2],
                        const void* const x[2], const double* const sb[2], const double cb[2], const int bf[2], const void* const y[2],
                        cudaStream_t s) {
  if (n <= 0) return;
  AxpbyPair P;
  bool al = true;
  for (int k = 0; k < 2; ++k) {
    P.out[k] = out[k]; P.x[k] = x[k]; P.y[k] = y[k]; P.sa[k] = sa[k]; P.sb[k] = sb[k];
    P.ca[k] = ca[k]; P.cb[k] = cb[k]; P.af[k] = af[k]; P.bf[k] = bf[k];
    al = al && aligned16(out[k]) && aligned16(x[k]) && aligned16(y[k]);
  }
  // what the launch writes, as one range (for the next bundle launch's "operator state before the wait" rule)
  const char* wlo = std::min((const char*)out[0], (const char*)out[1]);
  const size_t whi_off = (size_t)n * dsize(dtype);
  const char* whi = std::max((const char*)out[0], (const char*)out[1]) + whi_off;
  if (dtype == JETS_F32) {
    if (al) launch_pdl(axpby_pair_kernel<float, true>, grid_for(n / 4 + 1, 2), kThreads, s, wlo, whi - wlo, P, n);
    else launch_pdl(axpby_pair_kernel<float, false>, grid_for(n, 2), kThreads, s, wlo, whi - wlo, P, n);
  } else {
    if (al) launch_pdl(axpby_pair_kernel<double, true>, grid_for(n / 2 + 1, 2), kThreads, s, wlo, whi - wlo, P, n);
    else launch_pdl(axpby_pair_kernel<double, false>, grid_for(n, 2), kThreads, s, wlo, whi - wlo, P, n);
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

// ------------------------------------------------------------------ restriction ----------
// d = m[idx] and its adjoint m[idx] = d (zero elsewhere: the caller zero-fills first when it overwrites).
// Indices are unique (checked when the operator is built), so the scatter needs no atomics and is
// deterministic.  The index and the dense side stream coalesced; the sparse side goes through the
// read-only path.  HBM-bound: sizeof(I) + 2*sizeof(T) algorithmic bytes per selected element.
template <typename T, typename I>
__global__ void __launch_bounds__(kThreads) gather_kernel(T* __restrict__ out, const T* __restrict__ in,
                                                          const I* __restrict__ idx, int64_t n, int scatter, int acc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t j = (int64_t)idx[i];
    const int64_t o = scatter ? j : i, q = scatter ? i : j;
    const T v = in[q];
    if (acc == ACC_SET) out[o] = v;
    else if (acc == ACC_ADD) out[o] = out[o] + v;
    else out[o] = out[o] - v;
  }
}
template <typename T>
static void gather_t(void* out, const void* in, const void* idx, int idx64, int64_t n, int scatter, int acc, cudaStream_t s) {
  const unsigned g = grid_for(n, 4);
  if (idx64) gather_kernel<T, int64_t><<<g, kThreads, 0, s>>>((T*)out, (const T*)in, (const int64_t*)idx, n, scatter, acc);
  else gather_kernel<T, int32_t><<<g, kThreads, 0, s>>>((T*)out, (const T*)in, (const int32_t*)idx, n, scatter, acc);
}
void vec_gather(int dtype, void* out, const void* in, const void* idx, int idx64, int64_t n, int scatter, int acc, cudaStream_t s) {
  if (n <= 0) return;
  switch (dtype) {
    case JETS_F32: gather_t<float>(out, in, idx, idx64, n, scatter, acc, s); break;
    case JETS_F64: gather_t<double>(out, in, idx, idx64, n, scatter, acc, s); break;
    case JETS_C64: gather_t<Cx<float>>(out, in, idx, idx64, n, scatter, acc, s); break;
    default: gather_t<Cx<double>>(out, in, idx, idx64, n, scatter, acc, s); break;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void scalar_finish_norm(double*, double, cudaStream_t) {}

}  // namespace jets
