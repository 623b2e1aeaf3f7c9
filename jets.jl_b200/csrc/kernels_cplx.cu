// Vector-space kernels for the complex eltypes (ComplexF32 / ComplexF64 JetSpace, JetBSpace and the
// parent array of a JetSSpace): fill!, broadcast updates with complex coefficients, conj(x).*y,
// abs.(x), and the dot/norm reductions (src/Jets.jl:834-856 on complex blocks; test/runtests.jl
// :228-282, :542-550, :915-917).  `n` counts complex elements; storage is interleaved (re, im).
// Reductions are two-pass, fixed order, f64 accumulated -- no atomics, reproducible run to run --
// and optionally weighted: a SymmetricArray's norm runs over the LOGICAL array, where a stored
// element stands for 1 + (number of mirrored positions that map onto it) entries (:455-462).
#include "common.hpp"
#include "cplx.cuh"

namespace jets {
namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(int64_t n, int per_thread) {
  int64_t g = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  const int64_t cap = (int64_t)ctx().sm_count * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

template <typename R>
__global__ void __launch_bounds__(kThreads) cfill_kernel(Cx<R>* __restrict__ p, int64_t n, Cx<R> a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = a;
}

struct CLinArgs {
  const void* x[4];
  double re[4], im[4];
};
// out .= c0.*x0 .+ c1.*x1 ... left to right, one complex product and one sum per term
template <typename R, int K>
__global__ void __launch_bounds__(kThreads) clincomb_kernel(Cx<R>* __restrict__ out, int64_t n, CLinArgs a) {
  Cx<R> c[K];
#pragma unroll
  for (int k = 0; k < K; ++k) c[k] = Cx<R>((R)a.re[k], (R)a.im[k]);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Cx<R> r = c[0] * reinterpret_cast<const Cx<R>*>(a.x[0])[i];
#pragma unroll
    for (int k = 1; k < K; ++k) r = r + c[k] * reinterpret_cast<const Cx<R>*>(a.x[k])[i];
    out[i] = r;
  }
}

template <typename R>
__global__ void __launch_bounds__(kThreads) chadamard_kernel(Cx<R>* __restrict__ out, const Cx<R>* x,
                                                             const Cx<R>* y, int64_t n, int conj_x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Cx<R> a = x[i];
    if (conj_x) a = conj(a);
    out[i] = a * y[i];
  }
}

template <typename R>
__global__ void __launch_bounds__(kThreads) cabs_kernel(R* __restrict__ out, const Cx<R>* x, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Cx<R> a = x[i];
    out[i] = (R)hypot((double)a.re, (double)a.im);
  }
}

enum CKind : int { C_DOT_RE = 0, C_DOT_IM, C_SUMSQ, C_SUMABS, C_NNZ, C_MAXABS, C_MINABS, C_SUMPOW };

template <int KIND>
__device__ __forceinline__ double c_identity() {
  if (KIND == C_MINABS) return __longlong_as_double(0x7ff0000000000000LL);
  return 0.0;
}
template <int KIND>
__device__ __forceinline__ double c_combine(double a, double b) {
  if (KIND == C_MAXABS) return fmax(a, b);
  if (KIND == C_MINABS) return fmin(a, b);
  return a + b;
}
template <int KIND>
__device__ __forceinline__ double c_map(double xr, double xi, double yr, double yi, double w, double p) {
  switch (KIND) {
    case C_DOT_RE: return xr * yr + xi * yi;          // Re(conj(x) y)
    case C_DOT_IM: return xr * yi - xi * yr;          // Im(conj(x) y)
    case C_SUMSQ: return w * (xr * xr + xi * xi);
    case C_SUMABS: return w * hypot(xr, xi);
    case C_NNZ: return (xr != 0.0 || xi != 0.0) ? w : 0.0;
    case C_MAXABS: case C_MINABS: return hypot(xr, xi);
    default: return w * pow(hypot(xr, xi), p);
  }
}
template <int KIND>
__device__ __forceinline__ double c_block_reduce(double v) {
  __shared__ double sh[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = c_combine<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = c_identity<KIND>();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) r = c_combine<KIND>(r, sh[i]);
  }
  __syncthreads();
  return r;
}
// pass 1: CTA b reduces the contiguous chunk [b*chunk, (b+1)*chunk) -> partial[b]
template <typename R, int KIND>
__global__ void __launch_bounds__(kThreads) creduce_pass1(const Cx<R>* __restrict__ x, const Cx<R>* __restrict__ y,
                                                          const double* __restrict__ w, int64_t n, int64_t chunk,
                                                          double p, double* __restrict__ partial) {
  const int64_t b0 = (int64_t)blockIdx.x * chunk;
  int64_t b1 = b0 + chunk;
  if (b1 > n) b1 = n;
  double acc = c_identity<KIND>();
  for (int64_t i = b0 + threadIdx.x; i < b1; i += kThreads) {
    const Cx<R> a = x[i];
    Cx<R> b = a;
    if (KIND == C_DOT_RE || KIND == C_DOT_IM) b = y[i];
    const double wt = w ? w[i] : 1.0;
    acc = c_combine<KIND>(acc, c_map<KIND>((double)a.re, (double)a.im, (double)b.re, (double)b.im, wt, p));
  }
  acc = c_block_reduce<KIND>(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// the same pass over a REAL vector (a symmetric space with a real eltype, src/Jets.jl:408-441: conj is the identity,
// the weights still count the mirrored positions)
template <typename R, int KIND>
__global__ void __launch_bounds__(kThreads) rreduce_pass1(const R* __restrict__ x, const double* __restrict__ w, int64_t n,
                                                          int64_t chunk, double p, double* __restrict__ partial) {
  const int64_t b0 = (int64_t)blockIdx.x * chunk;
  int64_t b1 = b0 + chunk;
  if (b1 > n) b1 = n;
  double acc = c_identity<KIND>();
  for (int64_t i = b0 + threadIdx.x; i < b1; i += kThreads) {
    const double a = (double)x[i];
    acc = c_combine<KIND>(acc, c_map<KIND>(a, 0.0, a, 0.0, w ? w[i] : 1.0, p));
  }
  acc = c_block_reduce<KIND>(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
template <int KIND>
__global__ void __launch_bounds__(kThreads) creduce_pass2(const double* __restrict__ partial, int np, int finish,
                                                          double p, double* __restrict__ out) {
  double v = c_identity<KIND>();
  for (int i = threadIdx.x; i < np; i += kThreads) v = c_combine<KIND>(v, partial[i]);
  v = c_block_reduce<KIND>(v);
  if (threadIdx.x == 0) {
    if (finish == 1) v = sqrt(v);
    else if (finish == 2) v = pow(v, 1.0 / p);
    out[0] = v;
  }
}

template <typename R, int KIND>
void creduce_launch(const void* x, const void* y, const double* w, int64_t n, double p, int finish, double* dev_out,
                    cudaStream_t s) {
  Context& c = ctx();
  const int64_t quantum = (int64_t)kThreads * 4;
  int64_t nb = (n + quantum - 1) / quantum;
  const int64_t cap = (int64_t)c.sm_count * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  int64_t chunk = ((n + nb - 1) / nb + quantum - 1) / quantum * quantum;
  if (chunk < quantum) chunk = quantum;
  nb = n > 0 ? (n + chunk - 1) / chunk : 1;
  JETS_CHECK((size_t)nb <= c.dev_scratch_elems, JETS_ERR_INVALID, "reduction scratch too small");
  creduce_pass1<R, KIND><<<(unsigned)nb, kThreads, 0, s>>>((const Cx<R>*)x, (const Cx<R>*)y, w, n, chunk, p, c.dev_scratch);
  creduce_pass2<KIND><<<1, kThreads, 0, s>>>(c.dev_scratch, (int)nb, finish, p, dev_out);
  CUDA_TRY(cudaGetLastError());
  count_launch(2);
}
template <typename R, int KIND>
void rreduce_launch(const void* x, const double* w, int64_t n, double p, int finish, double* dev_out, cudaStream_t s) {
  Context& c = ctx();
  const int64_t quantum = (int64_t)kThreads * 4;
  int64_t nb = (n + quantum - 1) / quantum;
  const int64_t cap = (int64_t)c.sm_count * 8;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  int64_t chunk = ((n + nb - 1) / nb + quantum - 1) / quantum * quantum;
  if (chunk < quantum) chunk = quantum;
  nb = n > 0 ? (n + chunk - 1) / chunk : 1;
  JETS_CHECK((size_t)nb <= c.dev_scratch_elems, JETS_ERR_INVALID, "reduction scratch too small");
  rreduce_pass1<R, KIND><<<(unsigned)nb, kThreads, 0, s>>>((const R*)x, w, n, chunk, p, c.dev_scratch);
  creduce_pass2<KIND><<<1, kThreads, 0, s>>>(c.dev_scratch, (int)nb, finish, p, dev_out);
  CUDA_TRY(cudaGetLastError());
  count_launch(2);
}
template <typename R>
void rreduce_dispatch(int kind, const void* x, const double* w, int64_t n, double p, int finish, double* out, cudaStream_t s) {
  switch (kind) {
    case 2: rreduce_launch<R, C_SUMSQ>(x, w, n, p, finish, out, s); break;
    case 3: rreduce_launch<R, C_SUMABS>(x, w, n, p, finish, out, s); break;
    case 4: rreduce_launch<R, C_NNZ>(x, w, n, p, finish, out, s); break;
    case 5: rreduce_launch<R, C_MAXABS>(x, w, n, p, finish, out, s); break;
    case 6: rreduce_launch<R, C_MINABS>(x, w, n, p, finish, out, s); break;
    case 7: rreduce_launch<R, C_SUMPOW>(x, w, n, p, finish, out, s); break;
    default: JETS_FAIL(JETS_ERR_INVALID, "bad weighted reduction kind %d", kind);
  }
}

template <typename R>
void creduce_dispatch(int kind, const void* x, const void* y, const double* w, int64_t n, double p, int finish,
                      double* out, cudaStream_t s) {
  switch (kind) {
    case 0: creduce_launch<R, C_DOT_RE>(x, y, w, n, p, finish, out, s); break;
    case 1: creduce_launch<R, C_DOT_IM>(x, y, w, n, p, finish, out, s); break;
    case 2: creduce_launch<R, C_SUMSQ>(x, y, w, n, p, finish, out, s); break;
    case 3: creduce_launch<R, C_SUMABS>(x, y, w, n, p, finish, out, s); break;
    case 4: creduce_launch<R, C_NNZ>(x, y, w, n, p, finish, out, s); break;
    case 5: creduce_launch<R, C_MAXABS>(x, y, w, n, p, finish, out, s); break;
    case 6: creduce_launch<R, C_MINABS>(x, y, w, n, p, finish, out, s); break;
    case 7: creduce_launch<R, C_SUMPOW>(x, y, w, n, p, finish, out, s); break;
    default: JETS_FAIL(JETS_ERR_INVALID, "bad complex reduction kind %d", kind);
  }
}

template <typename R>
void clincomb_k(Cx<R>* out, int64_t n, int k, const CLinArgs& a, cudaStream_t s) {
  const unsigned g = grid_for(n, 4);
  switch (k) {
    case 1: clincomb_kernel<R, 1><<<g, kThreads, 0, s>>>(out, n, a); break;
    case 2: clincomb_kernel<R, 2><<<g, kThreads, 0, s>>>(out, n, a); break;
    case 3: clincomb_kernel<R, 3><<<g, kThreads, 0, s>>>(out, n, a); break;
    default: clincomb_kernel<R, 4><<<g, kThreads, 0, s>>>(out, n, a); break;
  }
}

}  // namespace

void cvec_fill(int dtype, void* p, int64_t n, double re, double im, cudaStream_t s) {
  if (n <= 0) return;
  if (dtype == JETS_C64) cfill_kernel<float><<<grid_for(n, 4), kThreads, 0, s>>>((Cx<float>*)p, n, Cx<float>((float)re, (float)im));
  else cfill_kernel<double><<<grid_for(n, 4), kThreads, 0, s>>>((Cx<double>*)p, n, Cx<double>(re, im));
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void cvec_lincomb(int dtype, void* out, int64_t n, int k, const double* c, const void* const* x, cudaStream_t s) {
  if (n <= 0) return;
  JETS_CHECK(k >= 1 && k <= 4, JETS_ERR_INVALID, "lincomb supports 1..4 terms, got %d", k);
  bool all_real = true;
  for (int i = 0; i < k; ++i) all_real = all_real && c[2 * i + 1] == 0.0;
  if (all_real) {
    // real coefficients scale both parts (Julia: a::Real * z::Complex): the real kernel over 2n parts
    double cr[4];
    for (int i = 0; i < k; ++i) cr[i] = c[2 * i];
    vec_lincomb(real_of(dtype), out, 2 * n, k, cr, x, s);
    return;
  }
  CLinArgs a{};
  for (int i = 0; i < k; ++i) { a.x[i] = x[i]; a.re[i] = c[2 * i]; a.im[i] = c[2 * i + 1]; }
  if (dtype == JETS_C64) clincomb_k<float>((Cx<float>*)out, n, k, a, s);
  else clincomb_k<double>((Cx<double>*)out, n, k, a, s);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void cvec_hadamard(int dtype, void* out, const void* x, const void* y, int64_t n, int conj_x, cudaStream_t s) {
  if (n <= 0) return;
  if (dtype == JETS_C64)
    chadamard_kernel<float><<<grid_for(n, 4), kThreads, 0, s>>>((Cx<float>*)out, (const Cx<float>*)x, (const Cx<float>*)y, n, conj_x);
  else
    chadamard_kernel<double><<<grid_for(n, 4), kThreads, 0, s>>>((Cx<double>*)out, (const Cx<double>*)x, (const Cx<double>*)y, n, conj_x);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void cvec_abs(int dtype, void* out_real, const void* x, int64_t n, cudaStream_t s) {
  if (n <= 0) return;
  if (dtype == JETS_C64) cabs_kernel<float><<<grid_for(n, 4), kThreads, 0, s>>>((float*)out_real, (const Cx<float>*)x, n);
  else cabs_kernel<double><<<grid_for(n, 4), kThreads, 0, s>>>((double*)out_real, (const Cx<double>*)x, n);
  CUDA_TRY(cudaGetLastError());
  count_launch();
}

void cvec_reduce(int dtype, int kind, const void* x, const void* y, const double* w, int64_t n, double p, int finish,
                 double* dev_out, cudaStream_t s) {
  if (dtype == JETS_F32) { rreduce_dispatch<float>(kind, x, w, n, p, finish, dev_out, s); return; }
  if (dtype == JETS_F64) { rreduce_dispatch<double>(kind, x, w, n, p, finish, dev_out, s); return; }
  if (dtype == JETS_C64) creduce_dispatch<float>(kind, x, y, w, n, p, finish, dev_out, s);
  else creduce_dispatch<double>(kind, x, y, w, n, p, finish, dev_out, s);
}

}  // namespace jets
