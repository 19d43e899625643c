/*
 * qlb200.h -- C ABI of libqlb200.so: the B200 (sm_100a) Householder-QL hot path behind
 * MatrixFactorizations.jl's own plugin surface.
 *
 * Every entry point is what a Julia `ccall` from the reference's layout-dispatch hooks would
 * bind (see INTEGRATION.md for the Julia stub of each).  Plain C types only: int64_t
 * dimensions (Julia BlasInt on 64-bit), column-major arrays, by-value scalars, no torch types,
 * no exceptions across the boundary.
 *
 * Return value (`info`, LAPACK convention extended):
 *      0           success
 *     -i           argument i (1-based, counting the handle as argument 1) is illegal; the Julia
 *                  shim maps it to ArgumentError / DimensionMismatch exactly where the reference
 *                  throws (src/ql.jl:326-328,356-358,494,508)
 *     >0, <1000    numerical status (index of a zero pivot in a triangular solve / UL: the
 *                  reference's SingularException / `info` field, src/ul.jl:54,102-107)
 *     1000+e       CUDA runtime error e (cudaError_t); text via qlb200_last_error()
 *
 * Pointer kinds: functions ending in `_dev` take DEVICE pointers and run asynchronously on the
 * handle's stream (no hidden copies, caller owns all buffers, results valid after
 * qlb200_sync()).  Functions without the suffix take HOST pointers, stage through the handle's
 * device workspace (H2D, compute, D2H) and return after the results are in host memory.
 *
 * There is NO CPU fallback anywhere in this library: without a CUDA device every call fails
 * with 1000+cudaErrorNoDevice (or similar).
 *
 * Packed layout (bit-compatible with the reference's `QL.factors` / `QL.tau`, src/ql.jl:37-45):
 * for column k (1-based, k >= n-min(m,n)+1) with mu = m-n+k: factors[1:mu-1,k] = v_k tail
 * (v_k[mu] = 1 implicit), factors[mu,k] = L diagonal, factors[mu+1:m,k] = L below the diagonal;
 * tau index k-n+min(m,n).  Q = H_n ... H_1 ("backward", LAPACK ?geqlf).
 */
#ifndef QLB200_H
#define QLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef struct qlb200_ctx *qlb200_handle;

/* ---- handle ------------------------------------------------------------------------------- */
/* Library ABI version (major*1000+minor). */
int qlb200_version(void);
/* Create a handle bound to CUDA device `device` (its own non-blocking stream + workspace).
 * Calls on one handle are serialized by the caller; different handles are thread-safe. */
int qlb200_create(qlb200_handle *h, int device);
int qlb200_destroy(qlb200_handle h);
/* Run subsequent calls on the caller's cudaStream_t (NULL = the CUDA default stream). */
int qlb200_set_stream(qlb200_handle h, void *cuda_stream);
/* Go back to the handle's own non-blocking stream. */
int qlb200_reset_stream(qlb200_handle h);
/* Block until everything queued on the handle's stream has finished. */
int qlb200_sync(qlb200_handle h);
/* Text of the last error recorded on this handle ("" if none).  Valid until the next call. */
const char *qlb200_last_error(qlb200_handle h);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t qlb200_launch_count(qlb200_handle h);
/* Upload rate (GB/s) the first chunk of the last large host-pointer ql! saw (plain launches only: a graph replay does not
 * measure); 0 before the first such call.  Diagnostic: ~53 GB/s on an idle host, 17-19 GB/s with 8 GPUs uploading at once. */
double qlb200_upload_rate(qlb200_handle h);
/* Page-lock (cudaHostRegister) / release a caller's host array.  Host-pointer entry points accept any host memory, but a
 * registered array is copied at full PCIe speed (about 5x the pageable rate) and a repeated large ql! on it is replayed as one
 * CUDA graph.  Registering twice / releasing an unregistered array is not an error.  The Julia shim calls these from
 * `QLB200.pin!(A)` / `QLB200.unpin!(A)`. */
int qlb200_host_register(qlb200_handle h, void *ptr, int64_t bytes);
int qlb200_host_unregister(qlb200_handle h, void *ptr);
/* With option "profile" = 1 the trailing-update / Q-apply DMMA GEMM launches are bracketed by CUDA
 * event pairs on the launching stream.  This call synchronizes, returns the summed device time (ms),
 * the summed algorithmic flops and the number of launches since the last collect, and resets. */
int qlb200_profile_collect(qlb200_handle h, double *ms_total, double *flops_total, int64_t *launches);
/* Tunables: "nb" (panel width, multiple of 16 in [16,256]; 0 = auto: 192 for look-ahead ql! with min(m,n) >= 14336, else 128),
 * "lookahead" (2 = default: ql!/ul!/complex ql! factor the next panel on a 16-SM green-context partition while the other SMs apply
 * the previous one, for large matrices; 1 = always; 0 = sequential schedule), "la_min", "la_pnarrow_min" (8192: while the next panel
 * starts at or right of this column the panel partition also applies the finished panel to the next panel's columns, so the update
 * partition runs nothing but the big GEMMs; 0 = never), "host_split" / "host_split_min" (host-pointer ql! of large
 * tall/square matrices, see DESIGN.md 6: 2 = default, the uploaded column chunks join the look-ahead factorization itself -- and a
 * repeated call on the same PINNED host matrix replays the whole upload | factor | download schedule as one CUDA graph; 1 = split
 * schedule; 0 = chunk-joining without look-ahead; "chunk_cols", "la_join_ramp", "la_join_chunks" tune the default schedule),
 * "fuse_chunk" (fused batched factor + apply: matrices per
 * chunk, 0 = whole batch), "gemm_cfg"
 * (-1 auto, 0..3 force a DMMA tile configuration), "trail_splits" (0 auto, 1..8 split-K of W = C'V), "strip_rpt"
 * (1/2/4 minimum rows per thread in the panel kernel; default 1 = as many CTAs of the cluster as the rows fill), "profile" (0/1, see above), "graph" (1: repeated ql! calls with
 * the same arguments are replayed as a CUDA graph), "tcache" (1: ql! keeps the panels' T factors for the following
 * lmul!/rmul!/ldiv! on the same factors -- device-pointer callers must not modify the factors in between),
 * "overlap_t" (1: the panel's T is built on a side stream while W = C'V runs), "ul_cluster" (1: cluster strip kernel
 * for UL), "stream_in" (1: host-pointer ql! uploads the matrix in column chunks while the right-hand panels are already
 * being factored; "chunk_cols", "join_step" tune it), "bvar" (batched kernel variants: 0 auto; n = 32: 2 always one kernel, 3 always the two-phase split; 32 < n <= 64: 6 warp-pair kernel, 7 one warp per matrix, 9 the literal ?geql2 kernels),
 * "pair_sync" (default 8; GEMMs with exactly two column tiles run each pair as a 2-CTA cluster that meets every this
 * many k-tiles so that the shared operand is read from DRAM once, 0 = off), "batch_chunk_mb" (default 32; host-pointer batched calls run as an upload | compute | download pipeline over chunks of
 * about this many megabytes, 0 = one chunk).
 * Diagnostic: with the environment variable QLB_LA_TRACE set, a look-ahead ql! brackets its phases with CUDA events, synchronizes at
 * the end and prints the per-iteration timeline (update partition idle / next-panel update / panel / trailing update) to stderr. */
int qlb200_set_option(qlb200_handle h, const char *name, int64_t value);

/* ---- single matrix: ql! -------------------------------------------------------------------
 * Replaces qlfactUnblocked!(A::StridedMatrix{<:BlasFloat}) = QL(LAPACK.geqlf!(A)...)
 * (reference src/ql.jl:72, reached from ql!/ql_layout! src/ql.jl:114-117).
 * A (m x n, leading dimension lda >= m) is overwritten with the packed factors; tau gets
 * min(m,n) entries.
 * Float64 only for the single-matrix path (Float32 single matrices are listed as "next" in DESIGN.md);
 * Any m, n; strips taller than 16384 rows (the capacity of the cluster-resident panel kernel) take a slower
 * global-memory panel path.
 * The host-pointer form streams the matrix in by column chunks (right to left) while the right-hand panels are already
 * being factored, and streams finished panels back while the rest is still being factored. */
int qlb200_dgeqlf(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb200_dgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau);

/* ---- single matrix: lmul!(Q,B), lmul!(Q',B), rmul!(C,Q), rmul!(C,Q') ------------------------
 * Replaces materialize!(::Lmul{<:QLPackedQLayout}) / {<:AdjQLPackedQLayout}
 * (reference src/ql.jl:321-347 / 350-377) and materialize!(::Rmul...) (src/ql.jl:381-406 /
 * 409-435).  LAPACK ?ormql argument order.  side 'L': C (m x n) <- op(Q) C, Q is m x m built
 * from the k = min(m, nA) reflectors stored in the LAST k columns of the m x nA factors; pass
 * A = pointer to the first of those k columns.  side 'R': C (m x n) <- C op(Q), Q is n x n.
 * trans 'N' or 'T'.  Returns -i on a dimension mismatch (reference: DimensionMismatch). */
int qlb200_dormql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k,
                  const double *A, int64_t lda, const double *tau, double *C, int64_t ldc);
int qlb200_dormql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k,
                      const double *dA, int64_t lda, const double *dtau, double *dC, int64_t ldc);

/* ---- single matrix: ldiv!(F::QL, B), square ------------------------------------------------
 * Replaces materialize!(::Ldiv{<:QLPackedLayout}) square branch (reference src/ql.jl:440-447,
 * 467): B (n x nrhs) <- L^{-1} (Q' B) with L = tril(factors).  The reference's tall/wide
 * branches are @test_broken (test/test_ql.jl:148) and are not offered.  The host-pointer form
 * returns i > 0 if L[i,i] == 0 (SingularException); the asynchronous _dev form cannot and
 * produces Inf/NaN like ?trsm. */
int qlb200_dqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                    const double *tau, double *B, int64_t ldb);
int qlb200_dqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const double *dA, int64_t lda,
                        const double *dtau, double *dB, int64_t ldb);

/* ---- tall least squares, min ||A x - b||_2 for m >= n through the QL factors.  The reference's own
 * rectangular branch of materialize!(::Ldiv{<:QLPackedLayout}) (src/ql.jl:448-486) is wrong and @test_broken
 * (test/test_ql.jl:148); this is the corrected tall case: B (m x nrhs) <- Q'B, rows m-n+1..m of it <- L^{-1} (those rows), and
 * the rows are rotated so that on return X = B[1:n, :] (n x nrhs) -- the rows the reference's `\` cuts out of the buffer
 * ldiv! worked on (src/ql.jl:496-501, _cut_B(X, 1:n)); LAPACK ?gels convention, same as qlb200_dqrsolve -- and B[n+1:m, :]
 * holds the residual's coordinates (its norm is the residual norm).
 * Returns 0, -i, 1000+cudaError, or i > 0 when L[i,i] == 0.  (The wide minimum-norm solve: qlb200_dqlminnorm below.) */
int qlb200_dqllstsq(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                    const double *tau, double *B, int64_t ldb);
int qlb200_dqllstsq_dev(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda,
                        const double *dtau, double *dB, int64_t ldb);
/* N2, wide half (m < n): the minimum-norm solution of A x = b from the QL factors of the wide A -- what the reference's branch at
 * src/ql.jl:448-483 is meant to compute (it is @test_broken, test/test_ql.jl:148; README.md:40 advertises `ql(A) \ b`).  A = Q L with the
 * trapezoid L = tril(factors, n-m): c = Q'b, RQ of the trapezoid L = [0 R] Z, y = R^{-1} c, x = Z' [0; y].  B is n x nrhs (ldb >= n):
 * rows 1..m hold b on entry, rows 1..n hold x on exit -- the buffer the reference's `\` allocates (src/ql.jl:496-501).
 * Returns 0, -i, 1000+cudaError, or i > 0 when R[i,i] == 0 (rank deficient). */
int qlb200_dqlminnorm(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau,
                      double *B, int64_t ldb);
int qlb200_dqlminnorm_dev(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda,
                          const double *dtau, double *dB, int64_t ldb);

/* ---- SURVEY.md 8(f) N1 + the Float32 half of A4: the other BlasFloat element types of the same hook -------------------
 * Reference src/ql.jl:72 `qlfactUnblocked!(A::StridedMatrix{T}) where T<:BlasFloat = QL(LAPACK.geqlf!(A)...)` is hit by
 * Float32, ComplexF64 and ComplexF32 as well (the element-type sweep of test/test_ql.jl:35-56); the Q-apply loops carry the
 * conj at src/ql.jl:336,366,368,400,426,429 and ldiv! is src/ql.jl:440-447,467.
 *   z: ComplexF64, interleaved (re, im) = Julia's Matrix{ComplexF64}; lda/ldc/ldb in COMPLEX elements; tau complex.  LAPACK
 *      zgeqlf / zunmql conventions (beta real, H = I - tau v v^H, trans is 'N' or 'C').  Panel: complex kernels; trailing
 *      update: the FP64 DMMA GEMMs on the real 2x2-block form of V and T (csrc/complex.cu).
 *   s: Float32, c: ComplexF32: widened on the device to the d / z path and narrowed again (csrc/mixed.cu states why).
 * Same return codes as the d forms (i > 0 from ?qlsolve: L[i,i] == 0, host-pointer form only). */
int qlb200_zgeqlf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau);
int qlb200_zgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau);
int qlb200_zunmql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                  const void *tau, void *C, int64_t ldc);
int qlb200_zunmql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc);
int qlb200_zqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const void *A, int64_t lda, const void *tau, void *B, int64_t ldb);
int qlb200_zqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const void *dA, int64_t lda, const void *dtau, void *dB,
                        int64_t ldb);
int qlb200_cgeqlf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau);
int qlb200_cgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau);
int qlb200_cunmql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                  const void *tau, void *C, int64_t ldc);
int qlb200_cunmql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc);
int qlb200_cqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const void *A, int64_t lda, const void *tau, void *B, int64_t ldb);
int qlb200_cqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const void *dA, int64_t lda, const void *dtau, void *dB,
                        int64_t ldb);
/* rq! / RQPackedQ for the same element types (reference src/rq.jl:401 -> LAPACK.gerqf!, :130-291 -> LAPACK.ormrq!):
 * ?gerqf(A) = (?geqlf(A^H))^H with identical tau; ?unmrq on the conjugate-transposed reflectors with the other op. */
int qlb200_zgerqf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau);
int qlb200_zgerqf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau);
int qlb200_zunmrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                  const void *tau, void *C, int64_t ldc);
int qlb200_zunmrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc);
int qlb200_cgerqf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau);
int qlb200_cgerqf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau);
int qlb200_cunmrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                  const void *tau, void *C, int64_t ldc);
int qlb200_cunmrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc);
int qlb200_sgerqf(qlb200_handle h, int64_t m, int64_t n, float *A, int64_t lda, float *tau);
int qlb200_sgerqf_dev(qlb200_handle h, int64_t m, int64_t n, float *dA, int64_t lda, float *dtau);
int qlb200_sormrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda,
                  const float *tau, float *C, int64_t ldc);
int qlb200_sormrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *dA, int64_t lda,
                      const float *dtau, float *dC, int64_t ldc);
int qlb200_sgeqlf(qlb200_handle h, int64_t m, int64_t n, float *A, int64_t lda, float *tau);
int qlb200_sgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, float *dA, int64_t lda, float *dtau);
int qlb200_sormql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda,
                  const float *tau, float *C, int64_t ldc);
int qlb200_sormql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *dA, int64_t lda,
                      const float *dtau, float *dC, int64_t ldc);
int qlb200_sqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const float *A, int64_t lda, const float *tau, float *B,
                    int64_t ldb);
int qlb200_sqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const float *dA, int64_t lda, const float *dtau, float *dB,
                        int64_t ldb);

/* ---- banded QL in BandedMatrices.jl storage ------------------------------------------------------
 * Replaces banded_ql!(L::BandedMatrix{Float64}) (reference ext/MatrixFactorizationsBandedMatricesExt.jl:18-36) behind
 * ql(::BandedMatrix) / ql!(::BandedMatrix) (ext:14-16; the caller widens the lower bandwidth to max(l, l+u+m-n) first, as ext:14
 * does).  D = bandeddata(L): (l+u+1) x n, column-major with leading dimension ldd, A[i,j] = D[u+i-j, j] (0-based); tau has min(m,n)
 * entries.  LinearAlgebra.reflector! conventions (a pivot with a zero or empty tail gets tau = 2, beta = -alpha).  One CTA walks the
 * n dependent column steps: latency-bound by construction (csrc/banded.cu) -- a completeness path, not a throughput path. */
int qlb200_dgbqlf(qlb200_handle h, int64_t m, int64_t n, int64_t l, int64_t u, double *D, int64_t ldd, double *tau);
int qlb200_dgbqlf_dev(qlb200_handle h, int64_t m, int64_t n, int64_t l, int64_t u, double *dD, int64_t ldd, double *dtau);

/* ---- RQ through index reversal -------------------------------------------------------------
 * Replaces rq!(A::StridedMatrix{<:BlasFloat}) = RQ(LAPACK.gerqf!(A)...) (reference
 * src/rq.jl:401): gerqf(A) = geqlf(A^T)^T with identical tau. */
int qlb200_dgerqf(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb200_dgerqf_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau);

/* RQ's packed Q (reference src/rq.jl:118-137 lmul!(Q::RQPackedQ,B) = LAPACK.ormrq!('L','N',...), :183-202
 * adjoint, :234-291 rmul!): LAPACK ?ormrq argument order.  A is k x nq (lda >= k): row i holds the vector
 * of H(i) as returned by gerqf in the LAST k rows of its factors (pass a pointer to the first of them).
 * Implemented as ?ormql on the transposed reflectors (Q_rq = Q_ql'). */
int qlb200_dormrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k,
                  const double *A, int64_t lda, const double *tau, double *C, int64_t ldc);
int qlb200_dormrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k,
                      const double *dA, int64_t lda, const double *dtau, double *dC, int64_t ldc);

/* ---- UL: reverse-order partially pivoted LU -------------------------------------------------------
 * Replaces ul!(A, pivot; check) = generic_ulfact!(A, pivot) (reference src/ul.jl:146-196): A[p,:] = U*L,
 * U unit upper / L lower packed in A, ipiv[k] (1-based, min(m,n) entries, Julia BlasInt = int64) the row
 * interchanged with row k, first maximum wins ties (src/ul.jl:162).  Only rows/columns 1..min(m,n) are
 * eliminated (src/ul.jl:155-191); interchanges run over all n columns.  pivot = 0 is Val(false).
 * *info_singular (device int32 for _dev) receives the reference's `info` field: 0, or the largest k with
 * a zero pivot (the Julia shim raises SingularException when check = true, src/ul.jl:193). */
int qlb200_dulfact(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, int64_t *ipiv,
                   int pivot, int32_t *info_singular);
int qlb200_dulfact_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, int64_t *dipiv,
                       int pivot, int32_t *dinfo_singular);
/* ldiv!(F::UL, B), square (reference src/ul.jl:363-393): rows of B permuted by ipiv in reverse order,
 * then U^{-1} (unit upper), then L^{-1}.  Host form returns i > 0 if L[i,i] == 0. */
int qlb200_dulsolve(qlb200_handle h, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                    const int64_t *ipiv, double *B, int64_t ldb);
int qlb200_dulsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const double *dA, int64_t lda,
                        const int64_t *dipiv, double *dB, int64_t ldb);

/* ---- the FP64 tensor-core GEMM of the trailing update, exposed for tests and benchmarks -------
 * C = alpha*op(A)*op(B) + beta*C, column-major device pointers, op 'N' or 'T' (BLAS dgemm order). */
int qlb200_dgemm_dev(qlb200_handle h, char transa, char transb, int64_t m, int64_t n, int64_t k, double alpha,
                     const double *dA, int64_t lda, const double *dB, int64_t ldb, double beta, double *dC,
                     int64_t ldc);

/* ---- QR through index reversal ----------------------------------------------
 * Replaces _qrfactUnblocked!(::AbstractColumnMajor, ::AbstractStridedLayout, A::AbstractMatrix{<:BlasFloat}, tau)
 *   = QR(LAPACK.geqrf!(A, tau)...) (reference src/qr.jl:72) behind qrunblocked!/qrunblocked (src/qr.jl:73-101).
 * With J the exchange matrix, geqrf(A) = J geqlf(J A J) J with tau reversed (the identity the reference tests at
 * test/test_ql.jl:7-13), so the QL kernels are reused unchanged; output = LAPACK ?geqrf layout (R on and above the
 * diagonal, reflector i below the diagonal of column i), i.e. QR.factors / QR.tau. */
int qlb200_dgeqrf(qlb200_handle h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb200_dgeqrf_dev(qlb200_handle h, int64_t m, int64_t n, double *dA, int64_t lda, double *dtau);
/* lmul!/rmul! with QRPackedQ or its adjoint (LAPACK ?ormqr semantics: A is nq x k, reflector i in column i): Q_qr = J Q_ql J. */
int qlb200_dormqr(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A,
                  int64_t lda, const double *tau, double *C, int64_t ldc);
int qlb200_dormqr_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *dA,
                      int64_t lda, const double *dtau, double *dC, int64_t ldc);

/* ldiv!(F::QR, B) for m >= n (`qrunblocked(A) \ b`, reference src/qr.jl:193-204): B (m x nrhs) is overwritten, X = B[1:n, :]
 * (exact solve for m == n, least squares for m > n, LAPACK ?gels layout); i > 0: R[i,i] == 0. */
int qlb200_dqrsolve(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                    const double *tau, double *B, int64_t ldb);
int qlb200_dqrsolve_dev(qlb200_handle h, int64_t m, int64_t n, int64_t nrhs, const double *dA, int64_t lda,
                        const double *dtau, double *dB, int64_t ldb);

/* ---- batched small square QL (new surface; per-matrix semantics = src/ql.jl:72) ------------
 * `batch` independent n x n matrices, 1 <= n <= 64 -- register-tile kernels for n <= 32 (8/16/32 exactly, other orders zero-padded
 * into the next tile), shared-memory + DMMA compact-WY kernels for 32 < n <= 64; matrix b starts at A + b*strideA,
 * column-major with leading dimension lda; tau of matrix b at tau + b*strideTau (n entries).
 * Reference-equivalent workload: `for i: ql!(view(A,:,:,i))`. */
int qlb200_dgeqlf_batched(qlb200_handle h, int64_t n, double *A, int64_t lda, int64_t strideA,
                          double *tau, int64_t strideTau, int64_t batch);
int qlb200_sgeqlf_batched(qlb200_handle h, int64_t n, float *A, int64_t lda, int64_t strideA,
                          float *tau, int64_t strideTau, int64_t batch);
int qlb200_dgeqlf_batched_dev(qlb200_handle h, int64_t n, double *dA, int64_t lda, int64_t strideA,
                              double *dtau, int64_t strideTau, int64_t batch);
int qlb200_sgeqlf_batched_dev(qlb200_handle h, int64_t n, float *dA, int64_t lda, int64_t strideA,
                              float *dtau, int64_t strideTau, int64_t batch);

/* Batched lmul!(Q,B) (trans 'N', src/ql.jl:321-347) / lmul!(Q',B) (trans 'T', src/ql.jl:350-377):
 * B_b (n x nrhs, leading dimension ldb, matrix stride strideB) <- op(Q_b) B_b.  1 <= n <= 64 like the
 * factorization (register kernels on 8/16/32/64-row tiles, other orders zero-padded into the next tile). */
int qlb200_dormql_batched(qlb200_handle h, char trans, int64_t n, int64_t nrhs, const double *A,
                          int64_t lda, int64_t strideA, const double *tau, int64_t strideTau,
                          double *B, int64_t ldb, int64_t strideB, int64_t batch);
int qlb200_sormql_batched(qlb200_handle h, char trans, int64_t n, int64_t nrhs, const float *A,
                          int64_t lda, int64_t strideA, const float *tau, int64_t strideTau,
                          float *B, int64_t ldb, int64_t strideB, int64_t batch);
int qlb200_dormql_batched_dev(qlb200_handle h, char trans, int64_t n, int64_t nrhs, const double *dA,
                              int64_t lda, int64_t strideA, const double *dtau, int64_t strideTau,
                              double *dB, int64_t ldb, int64_t strideB, int64_t batch);
int qlb200_sormql_batched_dev(qlb200_handle h, char trans, int64_t n, int64_t nrhs, const float *dA,
                              int64_t lda, int64_t strideA, const float *dtau, int64_t strideTau,
                              float *dB, int64_t ldb, int64_t strideB, int64_t batch);

/* Batched ldiv!(F::QL, B), square (src/ql.jl:440-447,467): B_b <- L_b^{-1} (Q_b' B_b).
 * `info_out` (may be NULL; `batch` int32 entries, device memory for _dev) receives per matrix
 * 0 or the 1-based index of a zero diagonal entry of L. */
int qlb200_dqlsolve_batched(qlb200_handle h, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                            int64_t strideA, const double *tau, int64_t strideTau, double *B,
                            int64_t ldb, int64_t strideB, int64_t batch, int32_t *info_out);
int qlb200_sqlsolve_batched(qlb200_handle h, int64_t n, int64_t nrhs, const float *A, int64_t lda,
                            int64_t strideA, const float *tau, int64_t strideTau, float *B,
                            int64_t ldb, int64_t strideB, int64_t batch, int32_t *info_out);
int qlb200_dqlsolve_batched_dev(qlb200_handle h, int64_t n, int64_t nrhs, const double *dA, int64_t lda,
                                int64_t strideA, const double *dtau, int64_t strideTau, double *dB,
                                int64_t ldb, int64_t strideB, int64_t batch, int32_t *dinfo_out);
int qlb200_sqlsolve_batched_dev(qlb200_handle h, int64_t n, int64_t nrhs, const float *dA, int64_t lda,
                                int64_t strideA, const float *dtau, int64_t strideTau, float *dB,
                                int64_t ldb, int64_t strideB, int64_t batch, int32_t *dinfo_out);

/* ---- batched, fused: ql! followed by lmul!(Q',B) / lmul!(Q,B) / ldiv!(F,B) on the factors just computed --------------
 * BASELINE.json configs[2] ("batched QL ... with lmul!(Q',B)") and [4] ("... + ldiv! on 8 RHS") as ONE call: A and tau come
 * back exactly as from q?geqlf_batched, B as from ?ormql_batched / ?qlsolve_batched (same kernels, same bits).  op: 'N' (Q B),
 * 'T' or 'C' (Q'B), 'S' (L^{-1} Q'B; info[b] = i > 0 if L_b[i,i] == 0, info may be null).  Host pointers: the factors cross PCIe once in
 * each direction instead of three times.  Device pointers: the two kernels back to back (option "fuse_chunk" = matrices per
 * L2-sized chunk, default 0 = off: measured slower, neither kernel is HBM-bound). */
int qlb200_dgeqlf_apply_batched(qlb200_handle h, char op, int64_t n, int64_t nrhs, double *A, int64_t lda, int64_t strideA,
                                double *tau, int64_t strideTau, double *B, int64_t ldb, int64_t strideB, int64_t batch,
                                int32_t *info);
int qlb200_sgeqlf_apply_batched(qlb200_handle h, char op, int64_t n, int64_t nrhs, float *A, int64_t lda, int64_t strideA,
                                float *tau, int64_t strideTau, float *B, int64_t ldb, int64_t strideB, int64_t batch,
                                int32_t *info);
int qlb200_dgeqlf_apply_batched_dev(qlb200_handle h, char op, int64_t n, int64_t nrhs, double *dA, int64_t lda, int64_t strideA,
                                    double *dtau, int64_t strideTau, double *dB, int64_t ldb, int64_t strideB, int64_t batch,
                                    int32_t *dinfo);
int qlb200_sgeqlf_apply_batched_dev(qlb200_handle h, char op, int64_t n, int64_t nrhs, float *dA, int64_t lda, int64_t strideA,
                                    float *dtau, int64_t strideTau, float *dB, int64_t ldb, int64_t strideB, int64_t batch,
                                    int32_t *dinfo);

/* ---- batched work sharded over several GPUs of one box (host pointers) ----------------------------
 * `hs[0..nh)` are handles created on the devices to use; handle g gets the contiguous matrix-index range
 * [g*batch/nh, (g+1)*batch/nh) (+ remainder on the first handles) and runs it on its own device from its
 * own host thread.  No cross-GPU communication; results land in the caller's arrays. */
int qlb200_dgeqlf_batched_mg(qlb200_handle *hs, int nh, int64_t n, double *A, int64_t lda, int64_t strideA,
                             double *tau, int64_t strideTau, int64_t batch);
int qlb200_sgeqlf_batched_mg(qlb200_handle *hs, int nh, int64_t n, float *A, int64_t lda, int64_t strideA,
                             float *tau, int64_t strideTau, int64_t batch);
int qlb200_dormql_batched_mg(qlb200_handle *hs, int nh, char trans, int64_t n, int64_t nrhs, const double *A,
                             int64_t lda, int64_t strideA, const double *tau, int64_t strideTau, double *B,
                             int64_t ldb, int64_t strideB, int64_t batch);
int qlb200_sormql_batched_mg(qlb200_handle *hs, int nh, char trans, int64_t n, int64_t nrhs, const float *A,
                             int64_t lda, int64_t strideA, const float *tau, int64_t strideTau, float *B,
                             int64_t ldb, int64_t strideB, int64_t batch);
int qlb200_dqlsolve_batched_mg(qlb200_handle *hs, int nh, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                               int64_t strideA, const double *tau, int64_t strideTau, double *B, int64_t ldb,
                               int64_t strideB, int64_t batch, int32_t *info_out);
int qlb200_sqlsolve_batched_mg(qlb200_handle *hs, int nh, int64_t n, int64_t nrhs, const float *A, int64_t lda,
                               int64_t strideA, const float *tau, int64_t strideTau, float *B, int64_t ldb,
                               int64_t strideB, int64_t batch, int32_t *info_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* QLB200_H */
