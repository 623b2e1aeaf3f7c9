// Per-vector interpreter of an elementwise/stencil chain ("term") -- shared by the TMA engine
// (operands staged in shared memory by cp.async.bulk) and the LDG engine (operands read with
// guarded global loads).  One IEEE rounding per arithmetic op, evaluated in the oracle's order;
// the translation unit is compiled with -fmad=false so nothing is contracted into an FMA and
// the device result is bit-identical to oracle/jets_oracle.py on these paths.
#pragma once
#include "common.hpp"
#include "cplx.cuh"

namespace jets {

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };
// complex spaces (LDG engine only): one 128-bit vector = 2 ComplexF32 or 1 ComplexF64
template <> struct VecOf<Cx<float>>  { using type = float4;  static constexpr int V = 2; };
template <> struct VecOf<Cx<double>> { using type = double2; static constexpr int V = 1; };

// HEAVY=false instantiations only know x^2 (the transcendental bodies, double-precision pow in
// particular, are hundreds of instructions and would evict the streaming loop from the I-cache).
template <typename T, bool HEAVY>
__device__ __forceinline__ T pw_phi(int fn, T x, T p) {
  if constexpr (!HEAVY) {
    return x * x;
  } else {
    switch (fn) {
      case JETS_PW_SQUARE: return x * x;
      case JETS_PW_POWER:  return pow(x, p);
      case JETS_PW_EXP:    return exp(x);
      case JETS_PW_SIN:    return sin(x);
      case JETS_PW_LOG:    return log(x);
      case JETS_PW_ATAN:   return atan(x);
      default:             return tanh(x);
    }
  }
}
template <typename T, bool HEAVY>
__device__ __forceinline__ T pw_dphi(int fn, T x, T p) {
  if constexpr (!HEAVY) {
    return T(2) * x;
  } else {
    switch (fn) {
      case JETS_PW_SQUARE: return T(2) * x;
      case JETS_PW_POWER:  return p * pow(x, p - T(1));
      case JETS_PW_EXP:    return exp(x);
      case JETS_PW_SIN:    return cos(x);
      case JETS_PW_LOG:    return T(1) / x;
      case JETS_PW_ATAN:   return T(1) / (T(1) + x * x);
      default: { T t = tanh(x); return T(1) - t * t; }
    }
  }
}

__device__ __forceinline__ CStage load_stage(const CStage* p) {
  // 16-byte uniform load; global tables go through the read-only (L1) path
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  CStage s;
  *reinterpret_cast<uint4*>(&s) = r;
  return s;
}

// Evaluates one term on NV windows at once.  val[i][w] holds the chain value at block-local
// position p0[i] - HL + w.  `ld(k, i, out)` loads the window of operand stream k for vector i.
// Positions outside [0, len) may hold garbage; every stencil masks them with a select, so
// garbage never reaches a valid position.
template <typename T, int HL, int HR, int NV, bool HEAVY, class Loader>
__device__ __forceinline__ void eval_term(const CStage* __restrict__ stages, int nstages,
                                          Loader& ld, const int64_t (&p0)[NV], int64_t len,
                                          T (&val)[NV][HL + VecOf<T>::V + HR]) {
  constexpr int W = HL + VecOf<T>::V + HR;
#pragma unroll
  for (int i = 0; i < NV; ++i) ld(0, i, val[i]);
  int sidx = 1;
  for (int s = 0; s < nstages; ++s) {
    const CStage st = load_stage(stages + s);
    switch (st.op) {
      case S_SCALE: {
        if constexpr (IsCx<T>::value) {   // real a times a complex vector scales both parts
          using R = typename IsCx<T>::real;
          const R c = (R)st.c0;
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int w = 0; w < W; ++w) val[i][w] = mulr(c, val[i][w]);
        } else {
          const T c = (T)st.c0;
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int w = 0; w < W; ++w) val[i][w] = c * val[i][w];
        }
      } break;
      case S_DIAG: {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          T b[W];
          ld(sidx, i, b);
          if constexpr (IsCx<T>::value) {
            if (st.fn & kConjFlag) {
#pragma unroll
              for (int w = 0; w < W; ++w) b[w] = conj(b[w]);
            }
          }
#pragma unroll
          for (int w = 0; w < W; ++w) val[i][w] = b[w] * val[i][w];
        }
        ++sidx;
      } break;
      case S_PW_F: {
        const T p = (T)st.c0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int w = 0; w < W; ++w) val[i][w] = pw_phi<T, HEAVY>(st.fn, val[i][w], p);
      } break;
      case S_PW_J: {
        const T p = (T)st.c0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          T b[W];
          ld(sidx, i, b);
          if constexpr (IsCx<T>::value) {   // adjoint of the Jacobian: conj(phi'(mo)) .* d
#pragma unroll
            for (int w = 0; w < W; ++w) {
              T j = pw_dphi<T, HEAVY>(st.fn & ~kConjFlag, b[w], p);
              if (st.fn & kConjFlag) j = conj(j);
              val[i][w] = j * val[i][w];
            }
          } else {
#pragma unroll
            for (int w = 0; w < W; ++w) val[i][w] = pw_dphi<T, HEAVY>(st.fn, b[w], p) * val[i][w];
          }
        }
        ++sidx;
      } break;
      case S_FDIFF: {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
#pragma unroll
          for (int w = 0; w + 1 < W; ++w) {
            const int64_t p = p0[i] - HL + w;
            val[i][w] = (p + 1 < len) ? (val[i][w + 1] - val[i][w]) : T(0);
          }
          val[i][W - 1] = T(0);
        }
      } break;
      case S_BDIFF: {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
#pragma unroll
          for (int w = W - 1; w >= 1; --w) {
            const int64_t p = p0[i] - HL + w;
            const T l = (p >= 1) ? val[i][w - 1] : T(0);
            const T r = (p + 1 < len) ? val[i][w] : T(0);
            val[i][w] = l - r;
          }
          val[i][0] = T(0);
        }
      } break;
      case S_LAP: {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          T t[W];
#pragma unroll
          for (int w = 0; w < W; ++w) t[w] = val[i][w];
#pragma unroll
          for (int w = 0; w < W; ++w) {
            const int64_t p = p0[i] - HL + w;
            const T l = (w >= 1 && p >= 1) ? t[w - 1] : T(0);
            const T r = (w + 1 < W && p + 1 < len) ? t[w + 1] : T(0);
            val[i][w] = (l - T(2) * t[w]) + r;
          }
        }
      } break;
      case S_NEG: {
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int w = 0; w < W; ++w) val[i][w] = -val[i][w];
      } break;
      case S_CSCALE: {  // complex a: (c0 of this stage) + i (c0 of the S_CIMAG stage that follows)
        if constexpr (IsCx<T>::value) {
          using R = typename IsCx<T>::real;
          const T c((R)st.c0, (R)load_stage(stages + s + 1).c0);
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int w = 0; w < W; ++w) val[i][w] = c * val[i][w];
        }
      } break;
      case S_CIMAG: break;
      default: {  // S_ZERO
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int w = 0; w < W; ++w) val[i][w] = T(0);
      } break;
    }
  }
}

}  // namespace jets

namespace jets {

// ------------------------------------------------------------------ fast paths -----------
// Straight-line evaluation of the common chains on one 128-bit vector, reading the staged tile
// directly.  `b0` points at this thread's first element of the term's first stream; stream k of the
// term lives `k * stride` bytes further.  `first`: the vector starts at the block's first element;
// `last`: index (relative to the vector) of the block's last element, so element j has a right
// neighbour inside the block iff j < last.  Operation order is the interpreter's, bit for bit.
template <typename T>
struct FastIO {
  using Vec = typename VecOf<T>::type;
  static constexpr int V = VecOf<T>::V;
  const char* b0;
  int stride;
  __device__ __forceinline__ void vec(int k, T (&x)[V]) const {
    const Vec v = *reinterpret_cast<const Vec*>(b0 + k * stride);
    const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) x[j] = vs[j];
  }
  __device__ __forceinline__ T at(int k, int j) const {  // element j (may be -1 or V: halo)
    return *reinterpret_cast<const T*>(b0 + k * stride + j * (int)sizeof(T));
  }
};

// Bundle engine: stream 0 (the term's input block) lives in the x ring, the state streams (k >= 1)
// in the state slot, `stride` bytes apart.
template <typename T>
struct FastIO2 {
  using Vec = typename VecOf<T>::type;
  static constexpr int V = VecOf<T>::V;
  const char* bin;   // this thread's vector of the input tile
  const char* bst;   // this thread's vector of the term's first state stream
  int stride;
  __device__ __forceinline__ const char* base(int k) const { return k == 0 ? bin : bst + (k - 1) * stride; }
  __device__ __forceinline__ void vec(int k, T (&x)[V]) const {
    const Vec v = *reinterpret_cast<const Vec*>(base(k));
    const T* vs = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int j = 0; j < V; ++j) x[j] = vs[j];
  }
  __device__ __forceinline__ T at(int k, int j) const {
    return *reinterpret_cast<const T*>(base(k) + j * (int)sizeof(T));
  }
};

// EDGE=false: the tile touches neither end of its block, so no element needs a boundary mask
// (`first` is false and `last` is past the vector): same arithmetic, fewer selects.
template <typename T, bool EDGE = true>
__device__ __forceinline__ void st_fdiff(const T (&x)[VecOf<T>::V], T xr, int last, T (&o)[VecOf<T>::V]) {
  constexpr int V = VecOf<T>::V;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const T nx = (j + 1 < V) ? x[(j + 1 < V) ? j + 1 : j] : xr;
    o[j] = (!EDGE || j < last) ? (nx - x[j]) : T(0);
  }
}
template <typename T, bool EDGE = true>
__device__ __forceinline__ void st_bdiff(const T (&x)[VecOf<T>::V], T xl, bool first, int last, T (&o)[VecOf<T>::V]) {
  constexpr int V = VecOf<T>::V;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const T pv = (j > 0) ? x[(j > 0) ? j - 1 : 0] : xl;
    const T l = (!EDGE || j > 0 || !first) ? pv : T(0);
    const T r = (!EDGE || j < last) ? x[j] : T(0);
    o[j] = l - r;
  }
}
template <typename T, bool EDGE = true>
__device__ __forceinline__ void st_lap(const T (&x)[VecOf<T>::V], T xl, T xr, bool first, int last, T (&o)[VecOf<T>::V]) {
  constexpr int V = VecOf<T>::V;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const T pv = (j > 0) ? x[(j > 0) ? j - 1 : 0] : xl;
    const T nx = (j + 1 < V) ? x[(j + 1 < V) ? j + 1 : j] : xr;
    const T l = (!EDGE || j > 0 || !first) ? pv : T(0);
    const T r = (!EDGE || j < last) ? nx : T(0);
    o[j] = (l - T(2) * x[j]) + r;
  }
}

// Returns false when `pattern` has no fast path (the caller falls back to the interpreter).
template <typename T, bool EDGE = true, class IO>
__device__ __forceinline__ bool eval_fast(int pattern, const IO& io, const CStage* stages,
                                          bool first, int last, T (&o)[VecOf<T>::V]) {
  constexpr int V = VecOf<T>::V;
  T x[V];
  switch (pattern) {
    case PAT_COPY: io.vec(0, o); return true;
    case PAT_DIAG: {
      T w[V];
      io.vec(0, x); io.vec(1, w);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = w[j] * x[j];
      return true;
    }
    case PAT_FDIFF: io.vec(0, x); st_fdiff<T, EDGE>(x, io.at(0, V), last, o); return true;
    case PAT_BDIFF: io.vec(0, x); st_bdiff<T, EDGE>(x, io.at(0, -1), first, last, o); return true;
    case PAT_LAP: io.vec(0, x); st_lap<T, EDGE>(x, io.at(0, -1), io.at(0, V), first, last, o); return true;
    case PAT_SCALE: {
      const T c = (T)load_stage(stages).c0;
      io.vec(0, x);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = c * x[j];
      return true;
    }
    case PAT_LAP_SCALE: case PAT_FDIFF_SCALE: case PAT_BDIFF_SCALE: {
      const T c = (T)load_stage(stages + 1).c0;
      T s[V];
      io.vec(0, x);
      if (pattern == PAT_LAP_SCALE) st_lap<T, EDGE>(x, io.at(0, -1), io.at(0, V), first, last, s);
      else if (pattern == PAT_FDIFF_SCALE) st_fdiff<T, EDGE>(x, io.at(0, V), last, s);
      else st_bdiff<T, EDGE>(x, io.at(0, -1), first, last, s);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = c * s[j];
      return true;
    }
    case PAT_SCALE_LAP: case PAT_SCALE_FDIFF: case PAT_SCALE_BDIFF: {
      const T c = (T)load_stage(stages).c0;
      T y[V];
      io.vec(0, x);
#pragma unroll
      for (int j = 0; j < V; ++j) y[j] = c * x[j];
      if (pattern == PAT_SCALE_LAP) st_lap<T, EDGE>(y, c * io.at(0, -1), c * io.at(0, V), first, last, o);
      else if (pattern == PAT_SCALE_FDIFF) st_fdiff<T, EDGE>(y, c * io.at(0, V), last, o);
      else st_bdiff<T, EDGE>(y, c * io.at(0, -1), first, last, o);
      return true;
    }
    case PAT_LAP_DIAG: case PAT_FDIFF_DIAG: case PAT_BDIFF_DIAG: {  // streams: x, w
      T w[V], s[V];
      io.vec(0, x); io.vec(1, w);
      if (pattern == PAT_LAP_DIAG) st_lap<T, EDGE>(x, io.at(0, -1), io.at(0, V), first, last, s);
      else if (pattern == PAT_FDIFF_DIAG) st_fdiff<T, EDGE>(x, io.at(0, V), last, s);
      else st_bdiff<T, EDGE>(x, io.at(0, -1), first, last, s);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = w[j] * s[j];
      return true;
    }
    case PAT_DIAG_LAP: case PAT_DIAG_FDIFF: case PAT_DIAG_BDIFF: {  // streams: x, w
      T w[V], y[V];
      io.vec(0, x); io.vec(1, w);
#pragma unroll
      for (int j = 0; j < V; ++j) y[j] = w[j] * x[j];
      if (pattern == PAT_DIAG_LAP) st_lap<T, EDGE>(y, io.at(1, -1) * io.at(0, -1), io.at(1, V) * io.at(0, V), first, last, o);
      else if (pattern == PAT_DIAG_FDIFF) st_fdiff<T, EDGE>(y, io.at(1, V) * io.at(0, V), last, o);
      else st_bdiff<T, EDGE>(y, io.at(1, -1) * io.at(0, -1), first, last, o);
      return true;
    }
    case PAT_J2: {
      T m[V];
      io.vec(0, x); io.vec(1, m);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = (T(2) * m[j]) * x[j];
      return true;
    }
    case PAT_SQUARE: {
      io.vec(0, x);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = x[j] * x[j];
      return true;
    }
    case PAT_J2_FDIFF_DIAG: {  // streams: x, mo, w
      T m[V], w[V], y[V], s[V];
      io.vec(0, x); io.vec(1, m); io.vec(2, w);
#pragma unroll
      for (int j = 0; j < V; ++j) y[j] = (T(2) * m[j]) * x[j];
      const T yr = (T(2) * io.at(1, V)) * io.at(0, V);
      st_fdiff<T, EDGE>(y, yr, last, s);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = w[j] * s[j];
      return true;
    }
    case PAT_DIAG_BDIFF_J2: {  // streams: x, w, mo
      T w[V], m[V], y[V], s[V];
      io.vec(0, x); io.vec(1, w); io.vec(2, m);
#pragma unroll
      for (int j = 0; j < V; ++j) y[j] = w[j] * x[j];
      const T yl = io.at(1, -1) * io.at(0, -1);
      st_bdiff<T, EDGE>(y, yl, first, last, s);
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = (T(2) * m[j]) * s[j];
      return true;
    }
    default: return false;
  }
}

// The same for NV vectors of one thread with ONE dispatch: the pattern is a compile-time constant
// inside every case, so each case is the straight-line body of eval_fast for that pattern repeated NV
// times (ncu on the bundle kernel: the per-vector indirect branch of `switch (pattern)` carried ~20%
// of all stall samples; dispatching once per term instead of once per vector halves them).  The four
// patterns of block-banded operators are tested first with plain (predictable, uniform) branches.
template <typename T, bool EDGE, int NV, class IO>
__device__ __forceinline__ bool eval_fast_n(int pattern, const IO (&io)[NV], const CStage* stages,
                                            const bool (&first)[NV], const int (&last)[NV],
                                            T (&o)[NV][VecOf<T>::V]) {
#define JETS_PAT_BODY(P)                                                               \
  {                                                                                    \
    _Pragma("unroll") for (int i = 0; i < NV; ++i)                                     \
        eval_fast<T, EDGE>(P, io[i], stages, first[i], last[i], o[i]);                 \
    return true;                                                                       \
  }
  if (pattern == PAT_DIAG) JETS_PAT_BODY(PAT_DIAG)
  if (pattern == PAT_LAP) JETS_PAT_BODY(PAT_LAP)
  if (pattern == PAT_FDIFF) JETS_PAT_BODY(PAT_FDIFF)
  if (pattern == PAT_BDIFF) JETS_PAT_BODY(PAT_BDIFF)
  switch (pattern) {
    case PAT_COPY: JETS_PAT_BODY(PAT_COPY)
    case PAT_SCALE: JETS_PAT_BODY(PAT_SCALE)
    case PAT_J2_FDIFF_DIAG: JETS_PAT_BODY(PAT_J2_FDIFF_DIAG)
    case PAT_DIAG_BDIFF_J2: JETS_PAT_BODY(PAT_DIAG_BDIFF_J2)
    case PAT_LAP_SCALE: JETS_PAT_BODY(PAT_LAP_SCALE)
    case PAT_FDIFF_SCALE: JETS_PAT_BODY(PAT_FDIFF_SCALE)
    case PAT_BDIFF_SCALE: JETS_PAT_BODY(PAT_BDIFF_SCALE)
    case PAT_J2: JETS_PAT_BODY(PAT_J2)
    case PAT_SQUARE: JETS_PAT_BODY(PAT_SQUARE)
    case PAT_SCALE_LAP: JETS_PAT_BODY(PAT_SCALE_LAP)
    case PAT_SCALE_FDIFF: JETS_PAT_BODY(PAT_SCALE_FDIFF)
    case PAT_SCALE_BDIFF: JETS_PAT_BODY(PAT_SCALE_BDIFF)
    case PAT_LAP_DIAG: JETS_PAT_BODY(PAT_LAP_DIAG)
    case PAT_FDIFF_DIAG: JETS_PAT_BODY(PAT_FDIFF_DIAG)
    case PAT_BDIFF_DIAG: JETS_PAT_BODY(PAT_BDIFF_DIAG)
    case PAT_DIAG_LAP: JETS_PAT_BODY(PAT_DIAG_LAP)
    case PAT_DIAG_FDIFF: JETS_PAT_BODY(PAT_DIAG_FDIFF)
    case PAT_DIAG_BDIFF: JETS_PAT_BODY(PAT_DIAG_BDIFF)
    default: return false;
  }
#undef JETS_PAT_BODY
}

}  // namespace jets
