/*
 * jets_b200.h -- C ABI of libjets_b200.so: the B200-native operator-application path of Jets.jl.
 *
 * Every entry point below is what a Julia `ccall` shim (julia/JetsB200.jl)
 * binds in place of the cited reference code (paths relative to the reference checkout,
 * ChevronETC/Jets.jl v1.4.1).  Plain pointers, sizes and opaque handles only; no C++ or torch
 * types; no exceptions cross the boundary: every call returns a jets_status and records a
 * thread-local message readable through jets_last_error().
 *
 * Threading contract (same as the reference, whose Jet is a mutable, unsynchronised struct,
 * src/Jets.jl:133): one host thread drives a context; calls are asynchronous on the context's
 * CUDA stream and ordered per stream; calls that return scalars to the host (jets_dot,
 * jets_norm, jets_extrema, jets_buf_download) synchronise that stream.
 *
 * There is NO CPU fallback: every compute entry point fails with JETS_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef JETS_B200_H
#define JETS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JETS_B200_ABI_VERSION 1

typedef struct jets_buf_s* jets_buf;   /* device vector: flat storage + block offset table   */
typedef struct jets_op_s*  jets_op;    /* operator tree node (a "Jet", src/Jets.jl:133-142)  */
typedef struct jets_scalar_s* jets_scalar; /* device-resident scalar (graph-capturable loops) */
typedef struct jets_dist_op_s* jets_dist_op; /* rank-local part of an operator partitioned over GPUs */

typedef enum {
  JETS_OK = 0,
  JETS_ERR_INVALID = 1,      /* null/dead handle, bad argument                              */
  JETS_ERR_SHAPE = 2,        /* space mismatch (the reference surfaces these from broadcast) */
  JETS_ERR_DTYPE = 3,
  JETS_ERR_CUDA = 4,         /* CUDA runtime error or no usable sm_100 device                */
  JETS_ERR_UNSUPPORTED = 5,  /* "not implemented" (src/Jets.jl:131)                          */
  JETS_ERR_NOT_LINEAR = 6,   /* adjoint / dot_product_test of a nonlinear op (:392, :1211)   */
  JETS_ERR_NO_POINT = 7,     /* Jacobian applied before point!/jacobian                      */
  JETS_ERR_NCCL = 8
} jets_status;

/* JETS_C64 / JETS_C128 = ComplexF32 / ComplexF64, interleaved (re, im) like Julia's Array{Complex{T}}
 * (complex JetSpace / JetBSpace / JetSSpace: test/runtests.jl:58-75, :228-282, :542-550, :915-917).
 * Lengths and offsets always count ELEMENTS of the buffer's eltype.                               */
typedef enum { JETS_F32 = 0, JETS_F64 = 1, JETS_C64 = 2, JETS_C128 = 3 } jets_dtype;

/* mul! dispatch (src/Jets.jl:390-392): F -> f!, DF -> df!, DFT -> df'! */
typedef enum { JETS_MODE_F = 0, JETS_MODE_DF = 1, JETS_MODE_DFT = 2 } jets_mode;

/* pointwise nonlinear registry: phi and its derivative (fixture JopBar test/runtests.jl:20-25,
 * doc example docs/src/index.md:110-113) */
typedef enum {
  JETS_PW_SQUARE = 0,  /* x*x       ; phi' = 2*x                  */
  JETS_PW_POWER  = 1,  /* x^p       ; phi' = p*x^(p-1)            */
  JETS_PW_EXP    = 2,
  JETS_PW_SIN    = 3,
  JETS_PW_TANH   = 4,
  JETS_PW_LOG    = 5,  /* log(x)    ; phi' = 1/x                  */
  JETS_PW_ATAN   = 6   /* atan(x)   ; phi' = 1/(1+x*x)            */
} jets_pw_fn;

/* stencils (no reference definition: JetPack.jl is un-vendored; semantics in DESIGN.md §4) */
typedef enum {
  JETS_ST_FDIFF = 0,   /* d[i] = m[i+1]-m[i], last row zero                                  */
  JETS_ST_LAP   = 1    /* d[i] = (m[i-1]-2m[i])+m[i+1], zero outside (self-adjoint)          */
} jets_stencil_kind;

/* ---------------------------------------------------------------- context ---------------- */
int  jets_abi_version(void);
/* Binds the calling process to `device` (one process per GPU) and creates the stream.        */
int  jets_init(int device);
int  jets_shutdown(void);
const char* jets_last_error(void);
/* Adopt an external cudaStream_t (e.g. torch's current stream) / read the active one.        */
int  jets_stream_set(void* cuda_stream);
void* jets_stream_get(void);
/* Fork/join onto one of two internal high-priority auxiliary streams, so that a communication call
 * overlaps the compute calls that follow: fork(a) makes the aux stream wait for everything issued
 * so far and directs subsequent calls to it; main() directs calls back to the main stream (the aux
 * work keeps running); join(a) makes the main stream wait for the aux stream.  One operator (and one
 * reduction) must not be in flight on two streams at once: its plan's scheduler counters, temporaries and
 * the context's reduction scratch are per plan / per context, not per stream.                    */
int  jets_stream_fork(int aux);
int  jets_stream_main(void);
int  jets_stream_join(int aux);
int  jets_sync(void);
/* Introspection used by tests and bench: total kernels launched by this library so far.       */
int64_t jets_launch_count(void);
/* Diagnostics (JETS_B200_TRACE=1 at jets_init): per-CTA timelines of the last 64 fused block-apply launches, 160 CTA
 * records of 8 %globaltimer stamps each (see csrc/kernels_fused_bundle.cu); returns the launches traced so far.   */
int64_t jets_debug_trace(uint64_t* host, int64_t capacity);
int  jets_device_sm_count(void);

/* ------------------------------------------------- device storage: JetSpace / JetBSpace ---- */
/* zeros(R)/Array(R) for R::JetSpace (nblocks==1) or R::JetBSpace (src/Jets.jl:105-108,
 * :922-924).  block_len[i] = length(space(R,i)).  Storage is ONE contiguous device buffer;
 * block i starts at element offset first(indices(R,i))-1, the reference's cumulative 1-based
 * ranges (:742-748) shifted to 0-based -- bit-exact block indexing.  Zero-initialised.        */
int jets_buf_create(jets_dtype dt, int32_t nblocks, const int64_t* block_len, jets_buf* out);
/* Wrap caller-owned device memory (e.g. a torch tensor): no copy, no guard padding.           */
int jets_buf_wrap(jets_dtype dt, void* devptr, int32_t nblocks, const int64_t* block_len, jets_buf* out);
/* getblock(x, i) (src/Jets.jl:914): a view sharing memory (first_block is 0-based).           */
int jets_buf_view(jets_buf x, int32_t first_block, int32_t nblocks, jets_buf* out);
/* reshape(x, R::JetBSpace) (src/Jets.jl:1112): same memory, new block table (lengths must sum
 * to length(x)).                                                                              */
int jets_buf_reshape(jets_buf x, int32_t nblocks, const int64_t* block_len, jets_buf* out);
int jets_buf_retain(jets_buf x);
int jets_buf_destroy(jets_buf x);
int jets_buf_dtype(jets_buf x);
int32_t jets_buf_nblocks(jets_buf x);
int64_t jets_buf_length(jets_buf x);
/* indices(x, i) (src/Jets.jl:780,858): 1-based inclusive range of 0-based block i.            */
int jets_buf_block_range(jets_buf x, int32_t block, int64_t* first1, int64_t* last1);
void* jets_buf_devptr(jets_buf x);
/* setblock!/getblock!/convert(Array,x) (src/Jets.jl:862-868, :915-916). block<0 = whole vector;
 * `count` elements of the buffer's dtype; host memory is not retained after return.           */
int jets_buf_upload(jets_buf x, int32_t block, const void* host, int64_t count);
int jets_buf_download(jets_buf x, int32_t block, void* host, int64_t count);
/* Asynchronous variants for pinned host memory (no stream sync).                              */
int jets_buf_upload_async(jets_buf x, int32_t block, const void* host, int64_t count);
int jets_buf_download_async(jets_buf x, int32_t block, void* host, int64_t count);
/* x[first .. first+count) (0-based element offset) <-> host, synchronous: scalar getindex /
 * setindex! (src/Jets.jl:819-832) and SymmetricArray element access (:455-484).                */
int jets_buf_write(jets_buf x, int64_t first, const void* host, int64_t count);
int jets_buf_read(jets_buf x, int64_t first, void* host, int64_t count);
int jets_buf_copy(jets_buf dst, jets_buf src);                   /* dst .= src                 */
int jets_buf_fill(jets_buf x, double a);                         /* fill!(x,a)  :880-885       */
int jets_buf_fill_c(jets_buf x, double re, double im);           /* the same with a complex a  */
/* rand(R)/randn(R): counter-based (Philox4x32-10) stream keyed by (seed, element index), so the
 * same (seed, logical index) gives the same value for any block/GPU partition.
 * dist 0: U[0,1)   dist 1: N(0,1)                                                             */
int jets_buf_rand(jets_buf x, uint64_t seed, uint64_t index_offset, int dist);

/* ------------------------------------------- BlockArray reductions and broadcast updates ---- */
/* dot(x,y) (src/Jets.jl:850-856); norm(x,p) (:834-848) with p in {2,1,0,+Inf,-Inf,other};
 * extrema(x) (:870-878).  Fixed-order two-pass reductions accumulated in f64; no atomics.     */
int jets_dot(jets_buf x, jets_buf y, double* out);               /* complex: the real part      */
int jets_norm(jets_buf x, double p, double* out);
int jets_extrema(jets_buf x, double* mn, double* mx);            /* real eltypes only           */
/* dot(x,y) = sum conj(x_i) y_i for the complex eltypes (conj on the FIRST argument as :853 /
 * LinearAlgebra.dot); out_re_im[0..1] = (real, imag).  Real vectors give imag = 0.              */
int jets_dot_c(jets_buf x, jets_buf y, double* out_re_im);
/* norm over the LOGICAL array of a SymmetricArray (src/Jets.jl:443-462): x is the stored parent
 * (complex), w (Float64, one weight per stored element) = 1 + number of mirrored positions whose
 * index map lands on that element.  p as jets_norm.                                             */
int jets_norm_weighted(jets_buf x, jets_buf w, double p, double* out);
/* out .= abs.(x) for complex x; out has the real eltype (test/runtests.jl:545-547).             */
int jets_abs(jets_buf out, jets_buf x);
/* out .= c[0].*x[0] .+ c[1].*x[1] ... (n<=4; left-to-right, one rounding per op, matching the
 * BlockArray broadcast copyto! src/Jets.jl:905-911).  out may alias any x[i].                  */
int jets_lincomb(jets_buf out, int32_t n, const double* c, const jets_buf* x);
/* The same with complex coefficients c_re_im[2i], c_re_im[2i+1] (complex eltypes only).         */
int jets_lincomb_c(jets_buf out, int32_t n, const double* c_re_im, const jets_buf* x);
/* out .= x .* y  (mask application in dot_product_test, src/Jets.jl:1215-1216).                */
int jets_hadamard(jets_buf out, jets_buf x, jets_buf y);

/* Device-resident scalars: the same reductions/updates without a host round trip, so a whole
 * CG/LSQR iteration can be captured in one CUDA graph.                          */
int jets_scalar_create(jets_scalar* out);
int jets_scalar_destroy(jets_scalar s);
int jets_scalar_set(jets_scalar s, double v);
int jets_scalar_get(jets_scalar s, double* v);                   /* synchronises               */
int jets_dot_dev(jets_buf x, jets_buf y, jets_scalar out);
int jets_norm_dev(jets_buf x, double p, jets_scalar out);
/* scalar arithmetic on device: out = a (op) b, op in '+','-','*','/' ; 'n' -> -a ; 's' sqrt(a);
 * 'h' -> hypot(a,b).                                                                           */
int jets_scalar_op(jets_scalar out, char op, jets_scalar a, jets_scalar b);
/* n scalar operations in one launch (out[i] = a[i] op[i] b[i], in order; b[i] may be null).     */
int jets_scalar_prog(int32_t n, const jets_scalar* out, const char* op, const jets_scalar* a, const jets_scalar* b);
/* out .= (sa? *sa : ca) .* x .+ (sb? *sb : cb) .* y ; a null scalar handle means "use the
 * constant"; negate flags fold a sign; inv flags use the reciprocal (x ./ beta).              */
int jets_axpby_dev(jets_buf out, jets_scalar sa, double ca, int a_flags, jets_buf x,
                   jets_scalar sb, double cb, int b_flags, jets_buf y);
#define JETS_COEF_NEG 1
#define JETS_COEF_INV 2
/* Two such updates of equal length in ONE pass -- out1 = a1.*x1 .+ b1.*y1 and out2 = a2.*x2 .+ b2.*y2 -- every
 * input element read before either output element is written, so outputs may alias inputs: CG's x += a p,
 * r -= a q and LSQR's x += (phi/rho) w, w = v - (theta/rho) w share one read of the common vector and one launch.
 * Bit-identical to two jets_axpby_dev calls issued in that order on non-overlapping updates.                    */
int jets_axpby_pair_dev(jets_buf out1, jets_scalar s1a, double c1a, int f1a, jets_buf x1, jets_scalar s1b, double c1b, int f1b, jets_buf y1,
                        jets_buf out2, jets_scalar s2a, double c2a, int f2a, jets_buf x2, jets_scalar s2b, double c2b, int f2b, jets_buf y2);
/* CUDA-graph capture of a sequence of library calls on the context stream.                    */
int jets_graph_begin(void);
int jets_graph_end(void** graph_exec_out);
int jets_graph_launch(void* graph_exec);
int jets_graph_destroy(void* graph_exec);

/* ----------------------------------------------------- primitive registry (leaf operators) -- */
/* JopLn(df! = d .= w.*m, df'! = m .= conj(w).*d) -- fixture JopFoo test/runtests.jl:3-8.      */
int jets_op_diag(jets_buf w, jets_op* out);
/* d .= a*m (_constdiag_df!, src/Jets.jl:1159-1160).                                            */
int jets_op_scale(jets_dtype dt, int64_t n, double a, jets_op* out);
/* complex a on a complex space; the adjoint applies conj(a) (_constdiag_df'!, :1160).          */
int jets_op_scale_c(jets_dtype dt, int64_t n, double a_re, double a_im, jets_op* out);
/* JopNl(f! = phi(m), df! = phi'(mo).*dm) -- fixture JopBar test/runtests.jl:20-25.            */
int jets_op_pointwise(jets_dtype dt, int64_t n, int fn, double p, jets_op* out);
int jets_op_stencil(jets_dtype dt, int64_t n, int kind, jets_op* out);
/* Matrix as operator (src/Jets.jl:325-326, :573-576; fixture JopBaz test/runtests.jl:27-33):
 * A is rows x cols, column-major (Julia layout), leading dimension = rows.  nrhs>1 applies A to
 * an (cols x nrhs) column-major matrix of right-hand sides (domain JetSpace(T,cols,nrhs)).
 * Complex eltypes: the adjoint applies the CONJUGATE transpose (mul!(m, A', d), :574).          */
int jets_op_dense(jets_buf A, int64_t rows, int64_t cols, int64_t nrhs, jets_op* out);
/* Restriction d = m[idx] with adjoint m[idx] = d, zero elsewhere -- the JetPack-style leaf the Jets
 * documentation composes with (docs/src/index.md:14-19; JetPack.jl itself is un-vendored, so the
 * definition is this library's: idx0 holds nidx UNIQUE 0-based positions of a length-n domain, copied
 * to the device).  Changes the length, so it is a fusion barrier with its own gather/scatter kernel;
 * plan_info bit6.                                                                               */
int jets_op_restrict(jets_dtype dt, int64_t n, int64_t nidx, const int64_t* idx0, jets_op* out);
/* JopZeroBlock(dom, rng) (src/Jets.jl:941-951).                                                */
int jets_op_zero(jets_dtype dt, int64_t ndom, int64_t nrng, jets_op* out);

/* ------------------------------------------------------------------------- combinators ------ */
/* JopLn(F)/JopLn(jet) view of a jet (src/Jets.jl:209-224): as a child it applies df! even in a
 * nonlinear parent.  Returns the op itself (retained) when already linear.                     */
int jets_op_as_linear(jets_op a, jets_op* out);
/* A' (src/Jets.jl:382-383); adjoint of an adjoint unwraps.  Fails with JETS_ERR_NOT_LINEAR for a
 * nonlinear op that is not wrapped by jets_op_as_linear.                                        */
int jets_op_adjoint(jets_op a, jets_op* out);
/* ops[0] ∘ ops[1] ∘ ... ∘ ops[n-1] (ops[n-1] is applied first), flattening nested composites
 * (src/Jets.jl:522-576).                                                                        */
int jets_op_compose(int32_t n, const jets_op* ops, jets_op* out);
/* sgn[0]*ops[0] + sgn[1]*ops[1] ... with sgn in {+1,-1}, flattening nested sums and flipping
 * their signs (src/Jets.jl:628-708).                                                            */
int jets_op_sum(int32_t n, const jets_op* ops, const int32_t* sgn, jets_op* out);
/* JopBlock / @blockop (src/Jets.jl:926-986): ops is nrow x ncol in COLUMN-MAJOR order (Julia's
 * Matrix layout).  dadom!=0 forces a block domain for a single column (:927).                   */
int jets_op_block(int32_t nrow, int32_t ncol, const jets_op* ops, int dadom, jets_op* out);
/* a*A (src/Jets.jl:1161-1164) = scale(range(A), a) ∘ A.  (The reference builds the scalar op on
 * domain(A), quirk Q6; identical for square A.)                                                 */
int jets_op_scalar_mul(double a, jets_op A, jets_op* out);
int jets_op_scalar_mul_c(double a_re, double a_im, jets_op A, jets_op* out);

/* ------------------------------------------------------------------ operator queries -------- */
int jets_op_retain(jets_op a);
int jets_op_destroy(jets_op a);                    /* Base.close (src/Jets.jl:290,1120)          */
int jets_op_is_linear(jets_op a);                  /* 1: JopLn/JopAdjoint-like, 0: JopNl-like    */
int jets_op_is_zero(jets_op a);                    /* iszero (:949-951)                          */
int jets_op_is_block(jets_op a);                   /* isblockop (:1097-1098)                     */
int jets_op_dtype(jets_op a);
/* nblocks(A, i) (:1074-1077) and block lengths of range (which=1) / domain (which=2).          */
int32_t jets_op_nblocks(jets_op a, int which);
int jets_op_block_len(jets_op a, int which, int32_t block, int64_t* len);
/* getblock(A, i, j) (:1085-1110), 0-based; returns a retained handle.                          */
int jets_op_getblock(jets_op a, int32_t i, int32_t j, jets_op* out);

/* ------------------------------------------------------------------ linearization ----------- */
/* point!(jet, mo) (src/Jets.jl:297-301; composite :578-589; sum :710-715; block :1059-1066).
 * The leaf stores mo BY REFERENCE (retains the buffer), as the reference does (:298).           */
int jets_op_set_point(jets_op a, jets_buf mo);
/* jacobian(F, mo) (src/Jets.jl:374): a NEW tree sharing the (immutable) state buffers, with a
 * private snapshot copy(mo); the result is linear.  jacobian!(F, mo) (:364-365) is
 * jets_op_set_point + jets_op_as_linear on the same jet.                                        */
int jets_op_jacobian(jets_op a, jets_buf mo, jets_op* out);
/* copy(A, false) (src/Jets.jl:230-233): a NEW tree of the same shape whose nodes share the (immutable)
 * state buffers and the current linearization points BY REFERENCE; a later point! on the copy does not
 * touch the original.  This is what `deepcopy(jet.s)` inside Jets' own `jacobian` (:374 -> :230) must do
 * with a device operator handle stored in the state.                                               */
int jets_op_clone(jets_op a, jets_op* out);

/* ------------------------------------------------------------------------- apply ------------ */
/* mul!(out, A, in) (src/Jets.jl:390-392).  accumulate!=0 reproduces quirk Q1: a
 * forward block apply with ncol>1 adds into `out` instead of overwriting it
 * (src/Jets.jl:1001,1024).  With accumulate==0 `out` is overwritten, which equals the
 * reference whenever `out` came from zeros(range(A)), i.e. for every `A*m` (:399).              */
int jets_apply(jets_op a, int mode, jets_buf out, jets_buf in, int accumulate);
/* out .= cA .* (A in) .+ cO .* out with device-resident coefficients (null scalar = use the
 * constant; JETS_COEF_* flags as for jets_axpby_dev): the Golub-Kahan updates u = A v - alpha u,
 * v = A'u - beta v of LSQR/CG (docs/src/index.md:235-246) in ONE pass -- fused into the store
 * epilogue of the block-apply kernel when the operator is elementwise/stencil, staged through a
 * temporary otherwise.                                                                          */
int jets_apply_axpby(jets_op a, int mode, jets_buf out, jets_buf in, jets_scalar sa, double ca, int a_flags,
                     jets_scalar so, double co, int o_flags);
/* The same, and norm2_out = norm(out, 2) of the vector just written (src/Jets.jl:834-848): LSQR's beta = ||u||,
 * alpha = ||v|| right after the Golub-Kahan updates, one call.  The norm runs directly behind the launch that wrote
 * the vector (for solver-sized vectors it is served from L2).  Folding the partial sums into the apply's store
 * epilogue was built and measured in round 2: neutral on config 4 and 8 % slower on every OTHER launch of the fused
 * kernel (instruction-issue pressure in its hot loop), so it was taken out again.                                */
int jets_apply_axpby_norm(jets_op a, int mode, jets_buf out, jets_buf in, jets_scalar sa, double ca, int a_flags,
                          jets_scalar so, double co, int o_flags, jets_scalar norm2_out);
/* Which engine the last plan for (op,mode) used: bit0 TMA-fused, bit1 LDG-fused, bit2 dense
 * GEMV, bit3 tcgen05 GEMM, bit4 staged through HBM temporaries, bit5 TMA-fused with the
 * shared-memory input-tile cache (rows sharing an input block fetch it once).;
 * bit6 gather/scatter (restriction).                                                          */
int jets_op_plan_info(jets_op a, int mode, int32_t* engines, int32_t* nlaunches);
/* Force an engine for A/B measurements: 0 auto, 1 TMA-fused, 2 LDG-fused, 3 TMA-fused without
 * the input-tile cache.                                                                         */
int jets_set_fused_engine(int which);

/* --------------------------------------------------------- multi-GPU (one process per GPU) --- */
/* Block-row partition of a JopBlock across ranks.  The 128-byte id is an
 * ncclUniqueId produced on rank 0 and broadcast by the host (torch.distributed / MPI / file).  */
int jets_dist_unique_id(char id[128]);
int jets_dist_init(int rank, int nranks, const char id[128]);
/* Host bootstrap, instead of or in addition to NCCL: `allgather(user, mine, all, bytes)` must gather `bytes`
 * bytes from every rank into all[rank*bytes ..] (MPI_Allgather, torch.distributed gloo, a shared file ...) and
 * return 0.  It carries only set-up records (CUDA IPC handles, layouts) and host scalars; the payload of the
 * jets_dist_op_* banded path moves through peer memory, so that path needs no NCCL at all.  The callback must
 * stay valid until jets_dist_shutdown.                                                                    */
typedef int (*jets_allgather_fn)(void* user, const void* mine, void* all, int64_t bytes);
int jets_dist_init_host(int rank, int nranks, jets_allgather_fn allgather, void* user);
int jets_dist_shutdown(void);
int jets_dist_rank(void);
int jets_dist_size(void);
/* Sum a host scalar over ranks in rank order (bit-stable dot/norm): all-gathers the partials. */
int jets_dist_sum_scalar(double* inout);
/* Dense-structure exchange on caller-owned vectors (what jets_dist_apply does inside for an operator made by
 * jets_dist_op_create_dense): all-gather equal domain shards / reduce-scatter per-rank partial domains (NCCL).  */
int jets_dist_allgather(jets_buf shard, jets_buf full);
int jets_dist_reduce_scatter(jets_buf full, jets_buf shard);


/* ------------------------------------------ distributed operators: ONE call per apply ---------- */
/* mul!(d, A, m) / mul!(m, A', d) (src/Jets.jl:390-392) for a JopBlock whose block rows are partitioned over
 * the ranks (one process per GPU).  The reference applies a block row as d_r = sum_c A_rc m_c
 * (src/Jets.jl:1015-1030) and a block column of the adjoint as m_c = sum_r A_rc' d_r (:1039-1055); here each
 * rank holds its rows and the matching shards of the vectors, and every exchange the sums need happens
 * INSIDE the call.
 *
 * Block-banded operator (A_rc == JopZeroBlock for |r-c| > halo): A_loc is the rank's nloc x (nloc+2*halo)
 * JopBlock over the halo-extended domain [halo blocks of rank-1 | nloc own blocks | halo blocks of rank+1]
 * (blocks outside the global operator are JopZeroBlock).  Collective: every rank creates its part in the same
 * order; neighbouring ranks map each other's exchange arena (CUDA IPC, same node).  Without jets_dist_init
 * (or halo == 0) the operator is purely local -- the same calls then serve one GPU.  The handle retains A_loc. */
int jets_dist_op_create(jets_op A_loc, int32_t halo, jets_dist_op* out);
/* Dense block structure: A_loc is the rank's nloc x ncol_total JopBlock over the WHOLE domain, which is
 * sharded over the ranks in equal contiguous pieces (domain length / nranks elements each).  Forward:
 * ncclAllGather of the shards, then the local rows; adjoint: local partial sums for the whole domain, then
 * ncclReduceScatter.  Needs jets_dist_init.                                                             */
int jets_dist_op_create_dense(jets_op A_loc, jets_dist_op* out);
int jets_dist_op_destroy(jets_dist_op A);
/* mode F/DF: out = rank-local rows of A*in, `in` = this rank's domain shard (its nloc own blocks; banded) or
 * domain/nranks elements (dense), `out` = its range blocks.  mode DFT: out = this rank's shard of A'*in,
 * `in` = its range blocks.  Banded operators: ONE kernel launch per rank per call -- the halo blocks are
 * written into the neighbours' arenas by the first work units of the launch (peer stores over NVLink), the
 * work units that need a neighbour's data run last and wait for its flag word; the adjoint's partial sums
 * are added in rank order (previous rank first), which for halo == 1 is bit-identical to the single-GPU
 * apply.  Asynchronous on the context stream; every rank must issue the same sequence of applies.        */
int jets_dist_apply(jets_dist_op A, int mode, jets_buf out, jets_buf in);
/* Registers a domain shard with the operator (collective: every rank registers its corresponding vector, in
 * the same order; library-owned vectors only).  The neighbours map the vector (CUDA IPC), and a FORWARD apply
 * whose `in` is a registered vector reads the halo blocks straight from the neighbours' memory (TMA loads over
 * NVLink inside the same single launch): no halo copy, the block rows that need a neighbour stay inside the
 * main row sweep, and the launch ends with a handshake (the neighbours are done reading) so that whatever
 * follows on the stream may overwrite `in`.  Unregistered vectors take the push path described above; the
 * results are bit-identical either way.                                                                   */
int jets_dist_op_register(jets_dist_op A, jets_buf x);
/* Host-buffer pipeline: host_out = A' * (A * host_in) for this rank's shards (nloc own blocks each, pinned
 * host memory), cut into `nchunks` block-row chunks (<= 0: default) pipelined on three internal streams:
 * upload k | forward k-1, adjoint k-2 | download k-2.  Bit-identical to upload, jets_dist_apply x2,
 * download.  Asynchronous: ordered after the context stream at call time; consecutive calls overlap (the
 * next upload runs under the previous download).  jets_dist_op_join makes the context stream wait for
 * everything issued so far (then jets_sync / events on the context stream see the results).              */
int jets_dist_apply_normal_host(jets_dist_op A, void* host_out, const void* host_in, int32_t nchunks);
int jets_dist_op_join(jets_dist_op A);
/* The compute-stream issue order of that pipeline, from the block structure alone (pure host function, no GPU:
 * the CPU tests replay it against NaN-poisoned buffers).  nz[r*(nloc+2*halo)+j] != 0 where block (r, j) of the
 * rank-local operator is not a zero block.  items receives (what, k) pairs -- what: 0 forward of chunk k, 1
 * adjoint of chunk k, 2 / 3 push the first / last halo blocks to the previous / next rank, 4 / 5 partial sums
 * for the previous / next rank's columns; chunk_bounds the [begin, end) block rows of every chunk; up_need[k]
 * the last upload chunk the forward of chunk k waits for.  Returns the number of items, -1 on error.        */
int32_t jets_dist_pipeline_schedule(int32_t nloc, int32_t halo, int32_t nchunks, int32_t has_prev, int32_t has_next,
                                    const uint8_t* nz, int32_t cap, int32_t* items, int32_t* chunk_bounds,
                                    int32_t* up_need, int32_t* nchunks_out);
/* what: 0 local block rows, 1 halo, 2 neighbours (bit0 previous, bit1 next), 3 chunks of the host pipeline,
 * 4 kernel launches of one jets_dist_apply, 5 kind (0 banded, 1 dense), 6 number of work units that gave up
 * waiting for a neighbour's flag (JETS_B200_GATE_TIMEOUT_MS, default 30 s; synchronises; results are invalid
 * when it is not 0 -- a rank that never issued the matching apply).                                       */
int32_t jets_dist_op_info(jets_dist_op A, int32_t what);

#ifdef __cplusplus
}
#endif
#endif /* JETS_B200_H */
