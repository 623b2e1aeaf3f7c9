// Per-vector interpreter of an elementwise/stencil chain ("term") -- shared by the TMA engine
// (operands staged in shared memory by cp.async.bulk) and the LDG engine (operands read with
// guarded global loads).  One IEEE rounding per arithmetic op, evaluated in the oracle's order;
// the translation unit is compiled with -fmad=false so nothing is contracted into an FMA and
// the device result is bit-identical to oracle/jets_oracle.py on these paths.
#pragma once
#include "common.hpp"

namespace jets {

template <typename T> struct VecOf;
template <> struct VecOf<float>  { using type = float4;  static constexpr int V = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int V = 2; };

// HEAVY=false instantiations only know x^2 (the transcendental bodies, double-precision pow in
// particular, are hundreds of instructions and would evict the streaming loop from the I-cache).
template <typename T, bool HEAVY>
__device__ __forceinline__ T pw_phi(int fn, T x, T p) {
  if (!HEAVY) return x * x;
  switch (fn) {
    case JETS_PW_SQUARE: return x * x;
    case JETS_PW_POWER:  return pow(x, p);
    case JETS_PW_EXP:    return exp(x);
    case JETS_PW_SIN:    return sin(x);
    default:             return tanh(x);
  }
}
template <typename T, bool HEAVY>
__device__ __forceinline__ T pw_dphi(int fn, T x, T p) {
  if (!HEAVY) return T(2) * x;
  switch (fn) {
    case JETS_PW_SQUARE: return T(2) * x;
    case JETS_PW_POWER:  return p * pow(x, p - T(1));
    case JETS_PW_EXP:    return exp(x);
    case JETS_PW_SIN:    return cos(x);
    default: { T t = tanh(x); return T(1) - t * t; }
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
        const T c = (T)st.c0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
          for (int w = 0; w < W; ++w) val[i][w] = c * val[i][w];
      } break;
      case S_DIAG: {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          T b[W];
          ld(sidx, i, b);
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
#pragma unroll
          for (int w = 0; w < W; ++w) val[i][w] = pw_dphi<T, HEAVY>(st.fn, b[w], p) * val[i][w];
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
