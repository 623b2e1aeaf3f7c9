/* C consumer of include/jets_b200.h: proves the header is plain C (compiled with -std=c99 -pedantic
 * -Wall -Werror) and that the calls a Julia `ccall` shim makes work from a non-Python, non-C++ host.
 *
 *   abi_smoke            -> exit 0 on a B200: builds the 2x2 JopBlock of diagonal blocks of
 *                           test/runtests.jl:622-666 (4 elements per block), applies it forward and
 *                           adjoint, checks both against the sums written out by hand, and runs the
 *                           dot-product identity  <A m, d> == <m, A' d>  (src/Jets.jl:1211-1226)
 *   abi_smoke (no GPU)   -> exit 3 after checking that jets_init fails with JETS_ERR_CUDA and a message
 *                           (there is no CPU fallback)
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "jets_b200.h"

#define CHECK(call)                                                                     \
  do {                                                                                  \
    int rc_ = (call);                                                                   \
    if (rc_ != JETS_OK) {                                                               \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, jets_last_error());           \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

int main(void) {
  enum { N = 4, NB = 2 };
  double w[NB][NB][N], m[NB * N], d[NB * N], out[NB * N], back[NB * N];
  int64_t lens[NB] = {N, N}, one = N, first1 = 0, last1 = 0;
  jets_buf wb[NB][NB], mb, db, ob, bb;
  jets_op leaf[NB * NB], A, At;
  double lhs, rhs;
  int i, r, c, rc;

  if (jets_abi_version() != JETS_B200_ABI_VERSION) {
    fprintf(stderr, "ABI version mismatch\n");
    return 1;
  }
  rc = jets_init(0);
  if (rc != JETS_OK) {
    if (rc == JETS_ERR_CUDA && strlen(jets_last_error()) > 0) {
      printf("no usable B200: jets_init -> JETS_ERR_CUDA (%s)\n", jets_last_error());
      return 3;
    }
    fprintf(stderr, "unexpected status %d from jets_init\n", rc);
    return 1;
  }
  for (r = 0; r < NB; ++r)
    for (c = 0; c < NB; ++c)
      for (i = 0; i < N; ++i) w[r][c][i] = 0.5 + 0.25 * r + 0.125 * c + 0.01 * i;
  for (i = 0; i < NB * N; ++i) {
    m[i] = 1.0 + 0.1 * i;
    d[i] = 2.0 - 0.05 * i;
  }
  for (r = 0; r < NB; ++r)
    for (c = 0; c < NB; ++c) {
      CHECK(jets_buf_create(JETS_F64, 1, &one, &wb[r][c]));
      CHECK(jets_buf_upload(wb[r][c], -1, w[r][c], N));
      CHECK(jets_op_diag(wb[r][c], &leaf[r + NB * c])); /* column-major, like Julia's Matrix{Jop} */
    }
  CHECK(jets_op_block(NB, NB, leaf, 0, &A));
  CHECK(jets_op_adjoint(A, &At));
  CHECK(jets_buf_create(JETS_F64, NB, lens, &mb));
  CHECK(jets_buf_create(JETS_F64, NB, lens, &db));
  CHECK(jets_buf_create(JETS_F64, NB, lens, &ob));
  CHECK(jets_buf_create(JETS_F64, NB, lens, &bb));
  CHECK(jets_buf_block_range(mb, 1, &first1, &last1));
  if (first1 != N + 1 || last1 != 2 * N) { /* indices(R, 2) == 5:8, src/Jets.jl:742-748 */
    fprintf(stderr, "block range (%lld, %lld)\n", (long long)first1, (long long)last1);
    return 1;
  }
  CHECK(jets_buf_upload(mb, -1, m, NB * N));
  CHECK(jets_buf_upload(db, -1, d, NB * N));
  CHECK(jets_apply(A, JETS_MODE_DF, ob, mb, 0));
  CHECK(jets_apply(At, JETS_MODE_DF, bb, db, 0));
  CHECK(jets_buf_download(ob, -1, out, NB * N));
  CHECK(jets_buf_download(bb, -1, back, NB * N));
  for (r = 0; r < NB; ++r)
    for (i = 0; i < N; ++i) {
      /* d_r = sum_c w_rc .* m_c, left to right (src/Jets.jl:1015-1030); m_c = sum_r w_rc .* d_r (:1039-1055) */
      double f = 0.0 + w[r][0][i] * m[i], t = 0.0 + w[0][r][i] * d[i];
      f = f + w[r][1][i] * m[N + i];
      t = t + w[1][r][i] * d[N + i];
      if (out[r * N + i] != f || back[r * N + i] != t) {
        fprintf(stderr, "mismatch at block %d element %d: %.17g vs %.17g / %.17g vs %.17g\n", r, i, out[r * N + i], f,
                back[r * N + i], t);
        return 1;
      }
    }
  CHECK(jets_dot(ob, db, &rhs));
  CHECK(jets_dot(mb, bb, &lhs));
  if (fabs(lhs - rhs) > 1e-12 * fabs(lhs + rhs)) {
    fprintf(stderr, "dot product test failed: %.17g vs %.17g\n", lhs, rhs);
    return 1;
  }
  printf("abi_smoke ok: 2x2 JopBlock forward/adjoint bit-exact, <m,A'd>=%.15g <Am,d>=%.15g, %lld kernel launches\n", lhs, rhs,
         (long long)jets_launch_count());
  CHECK(jets_op_destroy(At));
  CHECK(jets_op_destroy(A));
  for (i = 0; i < NB * NB; ++i) CHECK(jets_op_destroy(leaf[i]));
  for (r = 0; r < NB; ++r)
    for (c = 0; c < NB; ++c) CHECK(jets_buf_destroy(wb[r][c]));
  CHECK(jets_buf_destroy(mb));
  CHECK(jets_buf_destroy(db));
  CHECK(jets_buf_destroy(ob));
  CHECK(jets_buf_destroy(bb));
  CHECK(jets_shutdown());
  return 0;
}
