// mixed.cu -- Float32 and ComplexF32 single-matrix entry points (reference src/ql.jl:72 sends every BlasFloat down the same
// hook; test/test_ql.jl:35,42 sweep Float32/ComplexF32 too).  They are PROMOTED: the operands are widened on the device to
// Float64 / ComplexF64, run through the FP64 path (DMMA trailing update) and narrowed again.  Stated choice:
// B200 has no FP32 tensor-core mode that is exact (TF32 drops 13 mantissa bits), the FP64 DMMA pipe gives 24 TFLOP/s on ql!
// where an FFMA SGEMM-based path would peak near 2x that; results are at least as accurate as ?geqlf in working precision
// and meet the 50 n eps tolerance with eps = 2^-23 trivially.  A native FP32 panel + FFMA GEMM is the obvious next step.
#include "common.cuh"
#include "internal.h"

#define QLB_REQUIRE(h, cond, argno, msg)                      \
    do {                                                      \
        if (!(cond)) return qlb_fail(h, -(argno), msg "%s");  \
    } while (0)

namespace qlb {

template <typename S, typename D>
__global__ void cvt_kernel(const S *__restrict__ in, int64_t ldi, D *__restrict__ out, int64_t ldo, int64_t rows, int64_t cols) {
    const int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx % rows, c = idx / rows;
        out[r + c * ldo] = (D)in[r + c * ldi];
    }
}
template <typename S, typename D>
static int cvt(qlb200_ctx *h, const S *in, int64_t ldi, D *out, int64_t ldo, int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    int64_t blocks = (rows * cols + 255) / 256;
    if (blocks > (int64_t)h->sm_count * 16) blocks = (int64_t)h->sm_count * 16;
    cvt_kernel<S, D><<<(unsigned)blocks, 256, 0, h->stream>>>(in, ldi, out, ldo, rows, cols);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
static inline int64_t ev(int64_t x) { return x < 2 ? 2 : ((x + 1) & ~(int64_t)1); }

// MULT = 1: Float32 -> the d path; MULT = 2: ComplexF32 (interleaved) -> the z path.  All pointers are device pointers; leading
// dimensions in elements of the caller's type.
template <int MULT, bool RQ = false>
static int geqlf_promoted(qlb200_ctx *h, int64_t m, int64_t n, float *A, int64_t lda, float *tau) {
    const int64_t k = m < n ? m : n;
    if (k == 0) return 0;
    const int64_t rows = MULT * m, ldd = ev(rows);
    int rc = qlb_reserve(h, &h->cvt[0], &h->cvt_bytes[0], (size_t)ldd * n * sizeof(double));
    if (!rc) rc = qlb_reserve(h, &h->cvt[2], &h->cvt_bytes[2], (size_t)(MULT * k + 2) * sizeof(double));
    if (rc) return rc;
    double *dA = (double *)h->cvt[0], *dT = (double *)h->cvt[2];
    rc = cvt(h, A, MULT * lda, dA, ldd, rows, n);
    if (rc) return rc;
    if constexpr (RQ) rc = (MULT == 1) ? qlb_dgerqf_dev(h, m, n, dA, ldd, dT) : qlb_zgerqf_dev(h, m, n, dA, ldd / 2, dT);
    else rc = (MULT == 1) ? qlb_dgeqlf_dev(h, m, n, dA, ldd, dT) : qlb_zgeqlf_dev(h, m, n, dA, ldd / 2, dT);
    if (rc) return rc;
    rc = cvt(h, dA, ldd, A, MULT * lda, rows, n);
    if (rc) return rc;
    return cvt(h, dT, MULT * k, tau, MULT * k, MULT * k, 1);
}
// RQ: A is k x nq (rows = reflectors, LAPACK ?ormrq / ?unmrq), else nq x k (?ormql / ?unmql)
template <int MULT, bool RQ = false>
static int ormql_promoted(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda,
                          const float *tau, float *C, int64_t ldc) {
    if (m == 0 || n == 0 || k == 0) return 0;
    const int64_t nq = (side == 'L' || side == 'l') ? m : n;
    const int64_t arows = RQ ? k : nq, acols = RQ ? nq : k;
    const int64_t ra = MULT * arows, lda2 = ev(ra), rc_ = MULT * m, ldc2 = ev(rc_);
    int rc = qlb_reserve(h, &h->cvt[0], &h->cvt_bytes[0], (size_t)lda2 * acols * sizeof(double));
    if (!rc) rc = qlb_reserve(h, &h->cvt[1], &h->cvt_bytes[1], (size_t)ldc2 * n * sizeof(double));
    if (!rc) rc = qlb_reserve(h, &h->cvt[2], &h->cvt_bytes[2], (size_t)(MULT * k + 2) * sizeof(double));
    if (rc) return rc;
    double *dA = (double *)h->cvt[0], *dC = (double *)h->cvt[1], *dT = (double *)h->cvt[2];
    rc = cvt(h, A, MULT * lda, dA, lda2, ra, acols);
    if (!rc) rc = cvt(h, C, MULT * ldc, dC, ldc2, rc_, n);
    if (!rc) rc = cvt(h, tau, MULT * k, dT, MULT * k, MULT * k, 1);
    if (rc) return rc;
    if constexpr (RQ)
        rc = (MULT == 1) ? qlb_dormrq_dev(h, side, trans, m, n, k, dA, lda2, dT, dC, ldc2)
                         : qlb_zunmrq_dev(h, side, trans, m, n, k, dA, lda2 / 2, dT, dC, ldc2 / 2);
    else
        rc = (MULT == 1) ? qlb_dormql_dev(h, side, trans, m, n, k, dA, lda2, dT, dC, ldc2, -1)
                         : qlb_zunmql_dev(h, side, trans, m, n, k, dA, lda2 / 2, dT, dC, ldc2 / 2);
    if (rc) return rc;
    return cvt(h, dC, ldc2, C, MULT * ldc, rc_, n);
}
template <int MULT>
static int qlsolve_promoted(qlb200_ctx *h, int64_t n, int64_t nrhs, const float *A, int64_t lda, const float *tau, float *B,
                            int64_t ldb, int32_t *dinfo) {
    if (n == 0 || nrhs == 0) return 0;
    const int64_t ra = MULT * n, ld2 = ev(ra);
    int rc = qlb_reserve(h, &h->cvt[0], &h->cvt_bytes[0], (size_t)ld2 * n * sizeof(double));
    if (!rc) rc = qlb_reserve(h, &h->cvt[1], &h->cvt_bytes[1], (size_t)ld2 * nrhs * sizeof(double));
    if (!rc) rc = qlb_reserve(h, &h->cvt[2], &h->cvt_bytes[2], (size_t)(MULT * n + 2) * sizeof(double));
    if (rc) return rc;
    double *dA = (double *)h->cvt[0], *dB = (double *)h->cvt[1], *dT = (double *)h->cvt[2];
    rc = cvt(h, A, MULT * lda, dA, ld2, ra, n);
    if (!rc) rc = cvt(h, B, MULT * ldb, dB, ld2, ra, nrhs);
    if (!rc) rc = cvt(h, tau, MULT * n, dT, MULT * n, MULT * n, 1);
    if (rc) return rc;
    rc = (MULT == 1) ? qlb_dqlsolve_dev(h, n, nrhs, dA, ld2, dT, dB, ld2, dinfo, -1)
                     : qlb_zqlsolve_dev(h, n, nrhs, dA, ld2 / 2, dT, dB, ld2 / 2, dinfo);
    if (rc) return rc;
    return cvt(h, dB, ld2, B, MULT * ldb, ra, nrhs);
}

// host-pointer wrappers: the Float32 operands are staged as they are, widened on the device
static int up(qlb200_ctx *h, void **buf, size_t *have, const float *A, int64_t rows, int64_t cols, int64_t ld, size_t extra, float **d) {
    int rc = qlb_reserve(h, buf, have, ((size_t)(rows > 0 ? rows : 1) * (cols > 0 ? cols : 1) + extra) * sizeof(float));
    if (rc) return rc;
    *d = (float *)*buf;
    if (rows > 0 && cols > 0)
        QLB_CUDA(h, cudaMemcpy2DAsync(*d, (size_t)rows * 4, A, (size_t)ld * 4, (size_t)rows * 4, (size_t)cols, cudaMemcpyHostToDevice, h->stream));
    return 0;
}
static int down(qlb200_ctx *h, float *A, int64_t rows, int64_t cols, int64_t ld, const float *d) {
    if (rows > 0 && cols > 0)
        QLB_CUDA(h, cudaMemcpy2DAsync(A, (size_t)ld * 4, d, (size_t)rows * 4, (size_t)rows * 4, (size_t)cols, cudaMemcpyDeviceToHost, h->stream));
    return 0;
}

template <int MULT, bool RQ = false>
static int geqlf_any(qlb200_ctx *h, bool host, int64_t m, int64_t n, float *A, int64_t lda, float *tau) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 29), 2, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 3, "n out of range");
    QLB_REQUIRE(h, A != nullptr || m * n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (m > 1 ? m : 1), 5, "lda < max(1,m)");
    QLB_REQUIRE(h, tau != nullptr || m * n == 0, 6, "tau is null");
    const int64_t k = m < n ? m : n;
    if (k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    if (!host) return geqlf_promoted<MULT, RQ>(h, m, n, A, lda, tau);
    float *dA;
    int rc = up(h, &h->stage, &h->stage_bytes, A, MULT * m, n, MULT * lda, (size_t)MULT * k + 4, &dA);
    if (rc) return rc;
    float *dT = dA + (size_t)MULT * m * n;
    rc = geqlf_promoted<MULT, RQ>(h, m, n, dA, m, dT);
    if (rc) return rc;
    rc = down(h, A, MULT * m, n, MULT * lda, dA);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpyAsync(tau, dT, (size_t)MULT * k * 4, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
template <int MULT, bool RQ = false>
static int ormql_any(qlb200_ctx *h, bool host, char side, char trans, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda,
                     const float *tau, float *C, int64_t ldc) {
    if (!h) return -1;
    const bool left = side == 'L' || side == 'l', right = side == 'R' || side == 'r';
    QLB_REQUIRE(h, left || right, 2, "side must be 'L' or 'R'");
    const bool tn = trans == 'N' || trans == 'n', tc = trans == 'C' || trans == 'c', tt = trans == 'T' || trans == 't';
    QLB_REQUIRE(h, tn || tc || (tt && MULT == 1), 3, "trans must be 'N' or 'T' (real) / 'C' (complex)");
    QLB_REQUIRE(h, m >= 0 && m < (1 << 29), 4, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 5, "n out of range");
    const int64_t nq = left ? m : n;
    QLB_REQUIRE(h, k >= 0 && k <= nq, 6, "k must satisfy 0 <= k <= order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, A != nullptr || k == 0, 7, "A is null");
    const int64_t arows = RQ ? k : nq, acols = RQ ? nq : k;
    QLB_REQUIRE(h, lda >= (arows > 1 ? arows : 1), 8, "lda too small (DimensionMismatch)");
    QLB_REQUIRE(h, tau != nullptr || k == 0, 9, "tau is null");
    QLB_REQUIRE(h, C != nullptr || m * n == 0, 10, "C is null");
    QLB_REQUIRE(h, ldc >= (m > 1 ? m : 1), 11, "ldc < max(1,m)");
    if (m == 0 || n == 0 || k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    if (!host) return ormql_promoted<MULT, RQ>(h, side, trans, m, n, k, A, lda, tau, C, ldc);
    float *dA, *dC;
    int rc = up(h, &h->stage, &h->stage_bytes, A, MULT * arows, acols, MULT * lda, (size_t)MULT * k + 4, &dA);
    if (!rc) rc = up(h, &h->stage2, &h->stage2_bytes, C, MULT * m, n, MULT * ldc, 0, &dC);
    if (rc) return rc;
    float *dT = dA + (size_t)MULT * nq * k;
    QLB_CUDA(h, cudaMemcpyAsync(dT, tau, (size_t)MULT * k * 4, cudaMemcpyHostToDevice, h->stream));
    rc = ormql_promoted<MULT, RQ>(h, side, trans, m, n, k, dA, arows, dT, dC, m);
    if (rc) return rc;
    rc = down(h, C, MULT * m, n, MULT * ldc, dC);
    if (rc) return rc;
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
template <int MULT>
static int qlsolve_any(qlb200_ctx *h, bool host, int64_t n, int64_t nrhs, const float *A, int64_t lda, const float *tau, float *B, int64_t ldb) {
    if (!h) return -1;
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 2, "n out of range");
    QLB_REQUIRE(h, nrhs >= 0 && nrhs < (1 << 29), 3, "nrhs out of range");
    QLB_REQUIRE(h, A != nullptr || n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (n > 1 ? n : 1), 5, "lda < max(1,n)");
    QLB_REQUIRE(h, tau != nullptr || n == 0, 6, "tau is null");
    QLB_REQUIRE(h, B != nullptr || n * nrhs == 0, 7, "B is null");
    QLB_REQUIRE(h, ldb >= (n > 1 ? n : 1), 8, "ldb < max(1,n) (DimensionMismatch)");
    if (n == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    if (!host) return qlsolve_promoted<MULT>(h, n, nrhs, A, lda, tau, B, ldb, nullptr);
    float *dA, *dB;
    int rc = up(h, &h->stage, &h->stage_bytes, A, MULT * n, n, MULT * lda, (size_t)MULT * n + 4, &dA);
    if (!rc) rc = up(h, &h->stage2, &h->stage2_bytes, B, MULT * n, nrhs, MULT * ldb, 0, &dB);
    if (rc) return rc;
    float *dT = dA + (size_t)MULT * n * n;
    QLB_CUDA(h, cudaMemcpyAsync(dT, tau, (size_t)MULT * n * 4, cudaMemcpyHostToDevice, h->stream));
    rc = qlsolve_promoted<MULT>(h, n, nrhs, dA, n, dT, dB, n, h->dinfo);
    if (rc) return rc;
    rc = down(h, B, MULT * n, nrhs, MULT * ldb, dB);
    if (rc) return rc;
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return info == 0x7fffffff ? 0 : (int)info;
}

}  // namespace qlb
using namespace qlb;

extern "C" {
int qlb200_sgeqlf(qlb200_handle h, int64_t m, int64_t n, float *A, int64_t lda, float *tau) { return geqlf_any<1>(h, true, m, n, A, lda, tau); }
int qlb200_sgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, float *dA, int64_t lda, float *dtau) { return geqlf_any<1>(h, false, m, n, dA, lda, dtau); }
int qlb200_cgeqlf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau) { return geqlf_any<2>(h, true, m, n, (float *)A, lda, (float *)tau); }
int qlb200_cgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau) { return geqlf_any<2>(h, false, m, n, (float *)dA, lda, (float *)dtau); }
int qlb200_sgerqf(qlb200_handle h, int64_t m, int64_t n, float *A, int64_t lda, float *tau) { return geqlf_any<1, true>(h, true, m, n, A, lda, tau); }
int qlb200_sgerqf_dev(qlb200_handle h, int64_t m, int64_t n, float *dA, int64_t lda, float *dtau) { return geqlf_any<1, true>(h, false, m, n, dA, lda, dtau); }
int qlb200_cgerqf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau) { return geqlf_any<2, true>(h, true, m, n, (float *)A, lda, (float *)tau); }
int qlb200_cgerqf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau) { return geqlf_any<2, true>(h, false, m, n, (float *)dA, lda, (float *)dtau); }
int qlb200_sormrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda, const float *tau,
                  float *C, int64_t ldc) {
    return ormql_any<1, true>(h, true, side, trans, m, n, k, A, lda, tau, C, ldc);
}
int qlb200_sormrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *dA, int64_t lda,
                      const float *dtau, float *dC, int64_t ldc) {
    return ormql_any<1, true>(h, false, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
}
int qlb200_cunmrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda, const void *tau,
                  void *C, int64_t ldc) {
    return ormql_any<2, true>(h, true, side, trans, m, n, k, (const float *)A, lda, (const float *)tau, (float *)C, ldc);
}
int qlb200_cunmrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc) {
    return ormql_any<2, true>(h, false, side, trans, m, n, k, (const float *)dA, lda, (const float *)dtau, (float *)dC, ldc);
}
int qlb200_sormql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *A, int64_t lda, const float *tau,
                  float *C, int64_t ldc) {
    return ormql_any<1>(h, true, side, trans, m, n, k, A, lda, tau, C, ldc);
}
int qlb200_sormql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const float *dA, int64_t lda,
                      const float *dtau, float *dC, int64_t ldc) {
    return ormql_any<1>(h, false, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
}
int qlb200_cunmql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda, const void *tau,
                  void *C, int64_t ldc) {
    return ormql_any<2>(h, true, side, trans, m, n, k, (const float *)A, lda, (const float *)tau, (float *)C, ldc);
}
int qlb200_cunmql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc) {
    return ormql_any<2>(h, false, side, trans, m, n, k, (const float *)dA, lda, (const float *)dtau, (float *)dC, ldc);
}
int qlb200_sqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const float *A, int64_t lda, const float *tau, float *B, int64_t ldb) {
    return qlsolve_any<1>(h, true, n, nrhs, A, lda, tau, B, ldb);
}
int qlb200_sqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const float *dA, int64_t lda, const float *dtau, float *dB, int64_t ldb) {
    return qlsolve_any<1>(h, false, n, nrhs, dA, lda, dtau, dB, ldb);
}
int qlb200_cqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const void *A, int64_t lda, const void *tau, void *B, int64_t ldb) {
    return qlsolve_any<2>(h, true, n, nrhs, (const float *)A, lda, (const float *)tau, (float *)B, ldb);
}
int qlb200_cqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const void *dA, int64_t lda, const void *dtau, void *dB, int64_t ldb) {
    return qlsolve_any<2>(h, false, n, nrhs, (const float *)dA, lda, (const float *)dtau, (float *)dB, ldb);
}
}  // extern "C"
