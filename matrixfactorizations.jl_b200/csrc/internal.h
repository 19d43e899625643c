// internal.h -- typed entry points shared between the translation units of libqlb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
struct qlb200_ctx;

// carved from the handle's workspace by blocked.cu
struct QlbWorkspace {
    double *Vw;  int64_t ldv;      // clean reflector block, p x nb
    double *T;   int64_t ldt;      // nb x nb
    double *Ts;                    // 16 x 16 strip T
    double *G;                     // nb x nb Gram matrix
    double *Wsum, *W2; int64_t ldw; // (other dimension) x nb
    double *part; int64_t part_elems;
};
int qlb_block_workspace(qlb200_ctx *h, int64_t rows, int64_t other, int nb, QlbWorkspace *w);
// C <- (I - V Tm' V') C (left; Tm = T, or T' when tt) or C (I - V Tm V') (right): see apply_block in blocked.cu
int qlb_block_apply(qlb200_ctx *h, cudaStream_t st, const QlbWorkspace &ws, bool left, bool tt, const double *V, int64_t ldv,
                    const double *T, int64_t ldt, int p, int ib, double *C, int64_t ldc, int other);
int qlb_block_gram(qlb200_ctx *h, cudaStream_t st, const QlbWorkspace &ws, int p, int ib);

// complex.cu (device pointers, interleaved complex, asynchronous on h->stream)
int qlb_zgeqlf_dev(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb_zunmql_dev(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc);
int qlb_zgerqf_dev(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb_zunmrq_dev(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc);
int qlb_zqlsolve_dev(qlb200_ctx *h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                     int64_t ldb, int32_t *dinfo);

// batched.cu
int qlb_geqlf_batched_f64(qlb200_ctx *h, int64_t n, double *A, int64_t lda, int64_t strideA, double *tau,
                          int64_t strideTau, int64_t batch);
int qlb_geqlf_batched_f32(qlb200_ctx *h, int64_t n, float *A, int64_t lda, int64_t strideA, float *tau,
                          int64_t strideTau, int64_t batch);
int qlb_apply_batched_f64(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                          int64_t strideA, const double *tau, int64_t strideTau, double *B, int64_t ldb,
                          int64_t strideB, int64_t batch, int32_t *info_out);
int qlb_apply_batched_f32(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const float *A, int64_t lda,
                          int64_t strideA, const float *tau, int64_t strideTau, float *B, int64_t ldb,
                          int64_t strideB, int64_t batch, int32_t *info_out);

// gemm.cu: D = alpha*op(A)*op(B) + beta*D (DMMA).  akc/bkc: operand is K-contiguous (see gemm.cu).
int qlb_gemm(qlb200_ctx *h, cudaStream_t st, int cfg, bool akc, bool bkc, int M, int N, int K, double alpha,
             const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *D, int64_t ldd, int splits,
             int64_t split_stride);
int qlb_gemm_splits(int K, int splits, int bk);
int qlb_reduce_partials(qlb200_ctx *h, cudaStream_t st, const double *part, int S, int64_t stride, int M, int N,
                        int64_t ldp, double *out, int64_t ldo);

// panel.cu
int qlb_strip_max_rows();
int qlb_strip_factor(qlb200_ctx *h, cudaStream_t st, double *A, int64_t lda, int ps, int w, double *tau, double *Vw,
                     int64_t ldv, int p, double *T, int ldt);
int qlb_reduce_mult_t(qlb200_ctx *h, cudaStream_t st, const double *part, int S, int64_t stride, int rows, int w,
                      int64_t ldp, const double *T, int ldt, double *W2, int64_t ldw);
int qlb_larft(qlb200_ctx *h, cudaStream_t st, const double *G, int64_t ldg, const double *tau, int ib, double *T,
              int64_t ldt);
int qlb_extract_v(qlb200_ctx *h, cudaStream_t st, const double *F, int64_t ldf, int p, int ib, int mu0, double *Vw,
                  int64_t ldv);

// true if ql!(m x n) on device pointers takes the look-ahead path (two streams; not replayed as a graph)
bool qlb_lookahead_active(const qlb200_ctx *h, int64_t m, int64_t n);
int qlb_effective_nb(const qlb200_ctx *h, int64_t m, int64_t n);      // panel width ql!(m x n) uses (option "nb", 0 = auto)

// blocked.cu (device pointers, asynchronous on h->stream)
int qlb_dgeqlf_dev(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb_dormql_dev(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc, int cache_mode);
int qlb_dqlsolve_dev(qlb200_ctx *h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                     int64_t ldb, int32_t *dinfo, int cache_mode);
int qlb_dgerqf_dev(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb_dqlminnorm_dev(qlb200_ctx *h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                       int64_t ldb, int32_t *dinfo);
int qlb_dqrsolve_dev(qlb200_ctx *h, int64_t m, int64_t n, int64_t nrhs, const double *A, int64_t lda, const double *tau, double *B,
                     int64_t ldb, int32_t *dinfo);
int qlb_qr_flip_in(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda);
int qlb_qr_flip_out(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, double *tau);
int qlb_dormqr_dev(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc);
int qlb_rotate_rows_dev(qlb200_ctx *h, int64_t m, int64_t n, double *B, int64_t ldb, int64_t shift);
int trsm_grid(qlb200_ctx *h, int nrhs);
int qlb_trsm_lower_dev(qlb200_ctx *h, int n, int nrhs, const double *A, int64_t lda, double *B, int64_t ldb, int32_t *dinfo);
int qlb_dormrq_dev(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc);

// ul.cu (device pointers, asynchronous on h->stream)
int qlb_dulfact_dev(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, int64_t *ipiv, int pivot, int32_t *dinfo);
int qlb_dulsolve_dev(qlb200_ctx *h, int64_t n, int64_t nrhs, const double *A, int64_t lda, const int64_t *ipiv, double *B,
                     int64_t ldb, int32_t *dinfo);
