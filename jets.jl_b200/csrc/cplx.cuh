// Complex element type of the ComplexF32 / ComplexF64 spaces (JETS_C64 / JETS_C128): interleaved
// (re, im) storage, Julia's layout.  Arithmetic follows Julia Base's complex.jl: a product is
// (ar*br - ai*bi, ar*bi + ai*br) with one rounding per real operation (the translation units that
// use this header are compiled with -fmad=false), a real * complex product scales both parts.
#pragma once
#include <cuda_runtime.h>

namespace jets {

template <typename R>
struct Cx {
  R re, im;
  Cx() = default;
  __host__ __device__ explicit Cx(R r) : re(r), im(R(0)) {}
  __host__ __device__ Cx(R r, R i) : re(r), im(i) {}
};
template <typename R> __host__ __device__ __forceinline__ Cx<R> operator+(Cx<R> a, Cx<R> b) { return {a.re + b.re, a.im + b.im}; }
template <typename R> __host__ __device__ __forceinline__ Cx<R> operator-(Cx<R> a, Cx<R> b) { return {a.re - b.re, a.im - b.im}; }
template <typename R> __host__ __device__ __forceinline__ Cx<R> operator-(Cx<R> a) { return {-a.re, -a.im}; }
template <typename R> __host__ __device__ __forceinline__ Cx<R> operator*(Cx<R> a, Cx<R> b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename R> __host__ __device__ __forceinline__ Cx<R> mulr(R a, Cx<R> z) { return {a * z.re, a * z.im}; }
template <typename R> __host__ __device__ __forceinline__ Cx<R> conj(Cx<R> a) { return {a.re, -a.im}; }

template <typename T> struct IsCx { static constexpr bool value = false; using real = T; };
template <typename R> struct IsCx<Cx<R>> { static constexpr bool value = true; using real = R; };

// Stage flag (FStage::fn / CStage::fn bit 7): the operand stream enters conjugated -- the adjoint
// of a complex diagonal is m .= conj(w) .* d (fixture JopFoo, test/runtests.jl:3-8).  Real
// instantiations never see it.
constexpr int kConjFlag = 0x80;

}  // namespace jets
