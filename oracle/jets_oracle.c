/*
 * C restatement of the Jets.jl block-operator hot loops, used ONLY as (a) the CPU baseline that
 * bench.py times beside the GPU path and (b) a cross-check of the numpy oracle.
 * TEST / MEASUREMENT INFRASTRUCTURE -- never linked into or called from the product.
 * Parity status: unpinned (Julia is not installed; see oracle/jets_oracle.py header).
 *
 * Follows /root/reference/src/Jets.jl:
 *   JetBlock_df!   :1010-1032  (row-major traversal, dtmp temporary, `_d .+= mul!(dtmp, op, _m)`)
 *   JetBlock_df'!  :1034-1057  (column-major traversal, `_m .= 0`, mtmp, `_m .+= ...`)
 *   `A*m`          :399        (zeros(range(A)) allocation + fill before mul!)
 *   JopZeroBlock skip :1022,:1047
 * Leaves: diagonal (fixture JopFoo test/runtests.jl:3-4), and the build-defined stencils
 * (oracle/jets_oracle.py JopStencil).
 *
 * mode 0 "faithful": single thread, the reference's passes and temporaries (Jets has no threading).
 * mode 1 "threaded": the same arithmetic in the same order, fused per output element and spread over
 *                    all host cores with OpenMP -- the strongest CPU version of this path.
 * mode 2 "faithful-threaded": the reference's passes and temporaries (mode 0), every elementwise
 *                    pass spread over all host cores -- what Jets would do if Julia's broadcast
 *                    were multithreaded.
 * Compiled with -ffp-contract=off: one rounding per operation, like Julia's broadcast, so both
 * modes and the numpy oracle agree bit for bit.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { LEAF_ZERO = 0, LEAF_DIAG = 1, LEAF_FDIFF = 2, LEAF_LAP = 3 };

typedef struct {
  int32_t kind;
  int32_t pad;
  const void* state; /* diagonal weights */
} jets_ref_leaf;

int jets_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* The CPU arm of bench.py asks for every host core explicitly: launchers such as torchrun export
 * OMP_NUM_THREADS=1, which would silently turn the "all host threads" baseline into a one-thread one. */
void jets_ref_set_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

#define DEFINE(T, SUF)                                                                              \
  /* leaf applied to a whole block: out = op(in) (adj: op'(in)) */                                  \
  static void leaf_apply_##SUF(const jets_ref_leaf* op, int adj, T* out, const T* in, int64_t n,   \
                               int par) {                                                          \
    const T* w = (const T*)op->state;                                                               \
    int64_t i;                                                                                      \
    switch (op->kind) {                                                                             \
      case LEAF_DIAG:                                                                               \
        _Pragma("omp parallel for schedule(static) if(par)")                                        \
        for (i = 0; i < n; ++i) out[i] = w[i] * in[i];                                              \
        break;                                                                                      \
      case LEAF_FDIFF:                                                                              \
        if (!adj) {                                                                                 \
          _Pragma("omp parallel for schedule(static) if(par)")                                      \
          for (i = 0; i < n - 1; ++i) out[i] = in[i + 1] - in[i];                                   \
          out[n - 1] = 0;                                                                           \
        } else {                                                                                    \
          _Pragma("omp parallel for schedule(static) if(par)")                                      \
          for (i = 0; i < n; ++i) {                                                                 \
            const T l = i >= 1 ? in[i - 1] : (T)0;                                                  \
            const T r = i + 1 < n ? in[i] : (T)0;                                                   \
            out[i] = l - r;                                                                         \
          }                                                                                         \
        }                                                                                           \
        break;                                                                                      \
      case LEAF_LAP:                                                                                \
        _Pragma("omp parallel for schedule(static) if(par)")                                        \
        for (i = 0; i < n; ++i) {                                                                   \
          const T l = i >= 1 ? in[i - 1] : (T)0;                                                    \
          const T r = i + 1 < n ? in[i + 1] : (T)0;                                                 \
          out[i] = (l - (T)2 * in[i]) + r;                                                          \
        }                                                                                           \
        break;                                                                                      \
      default:                                                                                      \
        memset(out, 0, (size_t)n * sizeof(T));                                                      \
    }                                                                                               \
  }                                                                                                 \
  /* one output element of a leaf */                                                                \
  static inline T leaf_elem_##SUF(const jets_ref_leaf* op, int adj, const T* in, int64_t i,         \
                                  int64_t n) {                                                      \
    switch (op->kind) {                                                                             \
      case LEAF_DIAG: return ((const T*)op->state)[i] * in[i];                                      \
      case LEAF_FDIFF:                                                                              \
        if (!adj) return i + 1 < n ? in[i + 1] - in[i] : (T)0;                                      \
        else {                                                                                      \
          const T l = i >= 1 ? in[i - 1] : (T)0;                                                    \
          const T r = i + 1 < n ? in[i] : (T)0;                                                     \
          return l - r;                                                                             \
        }                                                                                           \
      case LEAF_LAP: {                                                                              \
        const T l = i >= 1 ? in[i - 1] : (T)0;                                                      \
        const T r = i + 1 < n ? in[i + 1] : (T)0;                                                   \
        return (l - (T)2 * in[i]) + r;                                                              \
      }                                                                                             \
      default: return (T)0;                                                                         \
    }                                                                                               \
  }                                                                                                 \
  /* out = A*in (adj=0) or A'*in (adj=1) for an R x C block operator, ops column-major.           \
     Square blocks only (elementwise leaves): dom_len[c] == rng_len[r] wherever op(r,c) != 0.  */  \
  int jets_ref_block_apply_##SUF(int32_t R, int32_t C, const jets_ref_leaf* ops,                    \
                                 const int64_t* dom_len, const int64_t* rng_len, const T* in,       \
                                 T* out, int adj, int mode) {                                       \
    const int32_t NO = adj ? C : R, NI = adj ? R : C;                                               \
    const int64_t* olen = adj ? dom_len : rng_len;                                                  \
    const int64_t* ilen = adj ? rng_len : dom_len;                                                  \
    int64_t* ooff = (int64_t*)malloc((size_t)(NO + 1) * sizeof(int64_t));                           \
    int64_t* ioff = (int64_t*)malloc((size_t)(NI + 1) * sizeof(int64_t));                           \
    int32_t o, k;                                                                                   \
    int64_t maxlen = 0;                                                                             \
    ooff[0] = 0;                                                                                    \
    for (o = 0; o < NO; ++o) { ooff[o + 1] = ooff[o] + olen[o]; if (olen[o] > maxlen) maxlen = olen[o]; } \
    ioff[0] = 0;                                                                                    \
    for (k = 0; k < NI; ++k) ioff[k + 1] = ioff[k] + ilen[k];                                       \
    if (mode == 0 || mode == 2) {                                                                   \
      const int par = (mode == 2);                                                                  \
      /* zeros(range(A)) of `A*m` (:399); fresh temporary like zeros(range(ops[1,1])) (:1012-1014) */ \
      T* tmp = (T*)calloc((size_t)(maxlen > 0 ? maxlen : 1), sizeof(T));                            \
      memset(out, 0, (size_t)ooff[NO] * sizeof(T));                                                 \
      for (o = 0; o < NO; ++o) {                                                                    \
        T* _o = out + ooff[o];                                                                      \
        const int64_t n = olen[o];                                                                  \
        int64_t i;                                                                                  \
        if (adj && NI > 1) memset(_o, 0, (size_t)n * sizeof(T)); /* _m .= 0 (:1042) */              \
        for (k = 0; k < NI; ++k) {                                                                  \
          const jets_ref_leaf* op = adj ? &ops[k + (size_t)o * R] : &ops[o + (size_t)k * R];        \
          if (op->kind == LEAF_ZERO) continue; /* iszero skip (:1022,:1047) */                      \
          if (NI > 1) {                                                                             \
            leaf_apply_##SUF(op, adj, tmp, in + ioff[k], n, par);                                   \
            _Pragma("omp parallel for schedule(static) if(par)")                                    \
            for (i = 0; i < n; ++i) _o[i] = _o[i] + tmp[i]; /* _d .+= dtmp (:1024,:1049) */         \
          } else {                                                                                  \
            leaf_apply_##SUF(op, adj, _o, in + ioff[k], n, par);                                    \
          }                                                                                         \
        }                                                                                           \
      }                                                                                             \
      free(tmp);                                                                                    \
    } else {                                                                                        \
      enum { CH = 4096 };                                                                           \
      /* compact per-output-block term lists (the iszero skip hoisted out of the hot loop) */       \
      int32_t* tptr = (int32_t*)malloc((size_t)(NO + 1) * sizeof(int32_t));                         \
      int32_t* tidx = (int32_t*)malloc((size_t)NO * (size_t)NI * sizeof(int32_t));                  \
      int32_t nt = 0;                                                                               \
      for (o = 0; o < NO; ++o) {                                                                    \
        tptr[o] = nt;                                                                               \
        for (k = 0; k < NI; ++k) {                                                                  \
          const jets_ref_leaf* op = adj ? &ops[k + (size_t)o * R] : &ops[o + (size_t)k * R];        \
          if (op->kind != LEAF_ZERO) tidx[nt++] = k;                                                \
        }                                                                                           \
      }                                                                                             \
      tptr[NO] = nt;                                                                                \
      _Pragma("omp parallel")                                                                       \
      {                                                                                             \
        int32_t ob;                                                                                 \
        for (ob = 0; ob < NO; ++ob) {                                                               \
          T* _o = out + ooff[ob];                                                                   \
          const int64_t n = olen[ob];                                                               \
          const int64_t nch = (n + CH - 1) / CH;                                                    \
          int64_t c;                                                                                \
          _Pragma("omp for schedule(static) nowait")                                                \
          for (c = 0; c < nch; ++c) {                                                               \
            T acc[CH];                                                                              \
            const int64_t i0 = c * CH, i1 = i0 + CH < n ? i0 + CH : n;                              \
            int64_t i;                                                                              \
            int32_t t;                                                                              \
            for (i = i0; i < i1; ++i) acc[i - i0] = 0;                                              \
            for (t = tptr[ob]; t < tptr[ob + 1]; ++t) {                                             \
              const int32_t kk = tidx[t];                                                           \
              const jets_ref_leaf* op = adj ? &ops[kk + (size_t)ob * R] : &ops[ob + (size_t)kk * R]; \
              const T* _in = in + ioff[kk];                                                         \
              if (op->kind == LEAF_DIAG) {                                                          \
                const T* w = (const T*)op->state;                                                   \
                for (i = i0; i < i1; ++i) acc[i - i0] = acc[i - i0] + w[i] * _in[i];                \
              } else {                                                                              \
                /* interior with branch-free (vectorisable) loops, block edges generically */       \
                const int64_t a0 = i0 > 1 ? i0 : 1, a1 = i1 < n - 1 ? i1 : n - 1;                   \
                if (i0 == 0) acc[0] = acc[0] + leaf_elem_##SUF(op, adj, _in, 0, n);                 \
                if (op->kind == LEAF_FDIFF && !adj) {                                               \
                  for (i = a0; i < a1; ++i) acc[i - i0] = acc[i - i0] + (_in[i + 1] - _in[i]);      \
                } else if (op->kind == LEAF_FDIFF) {                                                \
                  for (i = a0; i < a1; ++i) acc[i - i0] = acc[i - i0] + (_in[i - 1] - _in[i]);      \
                } else {                                                                            \
                  for (i = a0; i < a1; ++i)                                                         \
                    acc[i - i0] = acc[i - i0] + ((_in[i - 1] - (T)2 * _in[i]) + _in[i + 1]);        \
                }                                                                                   \
                if (i1 == n && n > 1)                                                               \
                  acc[n - 1 - i0] = acc[n - 1 - i0] + leaf_elem_##SUF(op, adj, _in, n - 1, n);      \
              }                                                                                     \
            }                                                                                       \
            for (i = i0; i < i1; ++i) _o[i] = acc[i - i0];                                          \
          }                                                                                         \
        }                                                                                           \
      }                                                                                             \
      free(tptr);                                                                                   \
      free(tidx);                                                                                   \
    }                                                                                               \
    free(ooff);                                                                                     \
    free(ioff);                                                                                     \
    return 0;                                                                                       \
  }                                                                                                 \
  /* config 2: d = w .* S(2*mo .* dm)  /  adjoint  m = (2*mo) .* S'(w .* d)  evaluated the way the  \
     reference's JetComposite_df!/df'! does (:530-540): one zero-filled temporary per stage + a     \
     final copy (mode 0), or fused and threaded (mode 1). */                                        \
  int jets_ref_chain_apply_##SUF(int64_t n, const T* w, const T* mo, const T* in, T* out, int adj,  \
                                 int mode) {                                                        \
    int64_t i;                                                                                      \
    if (mode == 0) {                                                                                \
      T* t1 = (T*)calloc((size_t)n, sizeof(T));                                                     \
      T* t2 = (T*)calloc((size_t)n, sizeof(T));                                                     \
      T* t3 = (T*)calloc((size_t)n, sizeof(T));                                                     \
      if (!adj) {                                                                                   \
        for (i = 0; i < n; ++i) t1[i] = ((T)2 * mo[i]) * in[i];                                     \
        for (i = 0; i + 1 < n; ++i) t2[i] = t1[i + 1] - t1[i];                                      \
        t2[n - 1] = 0;                                                                              \
        for (i = 0; i < n; ++i) t3[i] = w[i] * t2[i];                                               \
      } else {                                                                                      \
        for (i = 0; i < n; ++i) t1[i] = w[i] * in[i];                                               \
        for (i = 0; i < n; ++i) {                                                                   \
          const T l = i >= 1 ? t1[i - 1] : (T)0;                                                    \
          const T r = i + 1 < n ? t1[i] : (T)0;                                                     \
          t2[i] = l - r;                                                                            \
        }                                                                                           \
        for (i = 0; i < n; ++i) t3[i] = ((T)2 * mo[i]) * t2[i];                                     \
      }                                                                                             \
      memcpy(out, t3, (size_t)n * sizeof(T)); /* d .= dg(m) (:532) */                               \
      free(t1); free(t2); free(t3);                                                                 \
    } else if (!adj) {                                                                              \
      _Pragma("omp parallel for schedule(static)")                                                  \
      for (i = 0; i < n; ++i) {                                                                     \
        const T a = ((T)2 * mo[i]) * in[i];                                                         \
        const T s = i + 1 < n ? (((T)2 * mo[i + 1]) * in[i + 1]) - a : (T)0;                        \
        out[i] = w[i] * s;                                                                          \
      }                                                                                             \
    } else {                                                                                        \
      _Pragma("omp parallel for schedule(static)")                                                  \
      for (i = 0; i < n; ++i) {                                                                     \
        const T l = i >= 1 ? w[i - 1] * in[i - 1] : (T)0;                                           \
        const T r = i + 1 < n ? w[i] * in[i] : (T)0;                                                \
        out[i] = ((T)2 * mo[i]) * (l - r);                                                          \
      }                                                                                             \
    }                                                                                               \
    return 0;                                                                                       \
  }                                                                                                 \
  /* BlockArray dot / norm (:834-856): per-block partials combined in block order */               \
  double jets_ref_dot_##SUF(int32_t nb, const int64_t* len, const T* x, const T* y, int mode) {     \
    double a = 0;                                                                                   \
    int32_t b;                                                                                      \
    int64_t off = 0;                                                                                \
    for (b = 0; b < nb; ++b) {                                                                      \
      double s = 0;                                                                                 \
      int64_t i;                                                                                    \
      const T* xb = x + off;                                                                        \
      const T* yb = y + off;                                                                        \
      if (mode == 0) { for (i = 0; i < len[b]; ++i) s += (double)xb[i] * (double)yb[i]; }           \
      else {                                                                                        \
        _Pragma("omp parallel for reduction(+:s) schedule(static)")                                 \
        for (i = 0; i < len[b]; ++i) s += (double)xb[i] * (double)yb[i];                            \
      }                                                                                             \
      a += s;                                                                                       \
      off += len[b];                                                                                \
    }                                                                                               \
    return a;                                                                                       \
  }

DEFINE(float, f32)
DEFINE(double, f64)
