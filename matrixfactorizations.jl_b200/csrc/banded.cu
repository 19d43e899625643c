// banded.cu -- SURVEY.md 8(f) N4: Householder QL of a banded matrix in BandedMatrices.jl storage.
//
// Replaces banded_ql!(L::BandedMatrix) of the reference's BandedMatrices extension
// (ext/MatrixFactorizationsBandedMatricesExt.jl:18-36) for Float64; ql(A::BandedMatrix) (ext:14) widens the lower bandwidth to
// max(l, l+u+m-n) first and reaches it through ql!(A) (ext:16).  Band data D ((l+u+1) x n, column-major, ldd): A[i,j] = D[u+i-j, j]
// (0-based) = BandedMatrices.bandeddata.  For k = n ... max(n-m+2, 1) (1-based): reflector! on column k from its pivot row
// mu = m+k-n up to the top of the band, reflectorApply! to the columns j = max(1, mu-l) ... k-1.  Conventions =
// LinearAlgebra.reflector! (what the extension calls): tau = 0 only for an all-zero vector; a pivot with an empty or zero tail
// gets tau = 2, beta = -alpha (LAPACK ?larfg would leave it alone: SURVEY.md A.1).
//
// The algorithm is n strictly dependent steps of O((l+u)^2) flops each: one CTA walks the columns, a warp per updated column,
// the band in L2.  That is latency-bound on a GPU (a few microseconds per column, no matter how narrow the band) where a CPU
// core needs tens of nanoseconds for a narrow band -- it is offered for completeness of the drop-in (a BandedMatrix already on
// the device, wide bands), not as a throughput path; DESIGN.md section 8 says so.
#include "common.cuh"
#include "internal.h"

#define QLB_REQUIRE(h, cond, argno, msg)                      \
    do {                                                      \
        if (!(cond)) return qlb_fail(h, -(argno), msg "%s");  \
    } while (0)

namespace qlb {

constexpr int GB_THREADS = 256;

__global__ void __launch_bounds__(GB_THREADS, 1) gbqlf_kernel(double *__restrict__ D, int64_t ldd, int m, int n, int l, int u,
                                                              double *__restrict__ tau) {
    __shared__ double red[GB_THREADS / 32];
    __shared__ double sc[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = GB_THREADS / 32;
    const int kmin = (n - m + 2 > 1) ? n - m + 2 : 1;              // 1-based; real element type (ext:27)
    const int mn = m < n ? m : n;
    for (int i = tid; i < mn; i += GB_THREADS) tau[i] = 0.0;
    __syncthreads();
    for (int k = n; k >= kmin; --k) {                               // 1-based like the reference
        const int mu = m + k - n;
        const int d0 = u + 1 + mu - k;                              // data row of the pivot (1-based)
        const int dtop = (u - k + 2 > 1) ? u - k + 2 : 1;           // ... of the top of the band in column k
        const int N = d0 - dtop + 1;                                // length of the reflector (x[0] = pivot, going up)
        if (d0 < 1 || N <= 0) continue;                             // the pivot lies above the band: empty views
        double *xk = D + (int64_t)(k - 1) * ldd + (d0 - 1);         // x[r] = xk[-r]
        // ---- norm(x), scaled (LinearAlgebra.norm is overflow-safe) ----
        double am = 0.0;
        for (int r = tid; r < N; r += GB_THREADS) am = fmax(am, fabs(xk[-r]));
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
        if (lane == 0) red[warp] = am;
        __syncthreads();
        am = red[0];
        for (int w = 1; w < NW; ++w) am = fmax(am, red[w]);
        __syncthreads();
        if (am == 0.0) continue;                                    // normu == 0 -> tau = 0 (uniform over the block)
        double ss = 0.0;
        for (int r = tid; r < N; r += GB_THREADS) {
            const double q = xk[-r] / am;
            ss = fma(q, q, ss);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < NW; ++w) t += red[w];
            const double normu = am * sqrt(t);
            const double xi1 = xk[0];
            const double nu = copysign(normu, xi1);
            const double xi = xi1 + nu;
            sc[0] = xi;                 // divisor of the tail
            sc[1] = xi / nu;            // tau
            xk[0] = -nu;
            tau[k - n + mn - 1] = xi / nu;
        }
        __syncthreads();
        const double xi = sc[0], tk = sc[1];
        for (int r = 1 + tid; r < N; r += GB_THREADS) xk[-r] /= xi;
        __syncthreads();
        // ---- reflectorApply! to the columns j = max(1, mu - l) .. k-1: one warp per column ----
        const int j0 = (mu - l > 1) ? mu - l : 1;
        for (int j = j0 + warp; j < k; j += NW) {
            double *aj = D + (int64_t)(j - 1) * ldd + (u + 1 + mu - j - 1);      // a[r] = aj[-r]
            double v = 0.0;
            for (int r = 1 + lane; r < N; r += 32) v = fma(xk[-r], aj[-r], v);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            v = tk * (aj[0] + v);
            __syncwarp();
            if (lane == 0) aj[0] -= v;
            for (int r = 1 + lane; r < N; r += 32) aj[-r] = fma(-xk[-r], v, aj[-r]);
        }
        __syncthreads();
    }
}

}  // namespace qlb
using namespace qlb;

static int gb_check(qlb200_ctx *h, int64_t m, int64_t n, int64_t l, int64_t u, const double *D, int64_t ldd, const double *tau) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 30), 2, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 30), 3, "n out of range");
    QLB_REQUIRE(h, l >= 0 && u >= 0 && l + u + 1 < (1 << 24), 4, "bandwidths out of range");
    QLB_REQUIRE(h, l >= u + m - n || m * n == 0, 4, "lower bandwidth too small for the fill-in: widen to max(l, l+u+m-n) first (ext:14)");
    QLB_REQUIRE(h, D != nullptr || m * n == 0, 6, "band data is null");
    QLB_REQUIRE(h, ldd >= l + u + 1, 7, "ldd < l+u+1");
    QLB_REQUIRE(h, tau != nullptr || m * n == 0, 8, "tau is null");
    return 0;
}

extern "C" {

int qlb200_dgbqlf_dev(qlb200_handle h, int64_t m, int64_t n, int64_t l, int64_t u, double *dD, int64_t ldd, double *dtau) {
    int rc = gb_check(h, m, n, l, u, dD, ldd, dtau);
    if (rc) return rc;
    if (m == 0 || n == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    gbqlf_kernel<<<1, GB_THREADS, 0, h->stream>>>(dD, ldd, (int)m, (int)n, (int)l, (int)u, dtau);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
int qlb200_dgbqlf(qlb200_handle h, int64_t m, int64_t n, int64_t l, int64_t u, double *D, int64_t ldd, double *tau) {
    int rc = gb_check(h, m, n, l, u, D, ldd, tau);
    if (rc) return rc;
    if (m == 0 || n == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    const int64_t w = l + u + 1, k = m < n ? m : n;
    rc = qlb_reserve(h, &h->stage, &h->stage_bytes, ((size_t)w * n + (size_t)k) * sizeof(double));
    if (rc) return rc;
    double *dD = (double *)h->stage, *dtau = dD + (size_t)w * n;
    QLB_CUDA(h, cudaMemcpy2DAsync(dD, (size_t)w * 8, D, (size_t)ldd * 8, (size_t)w * 8, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    gbqlf_kernel<<<1, GB_THREADS, 0, h->stream>>>(dD, w, (int)m, (int)n, (int)l, (int)u, dtau);
    QLB_LAUNCH_CHECK(h);
    QLB_CUDA(h, cudaMemcpy2DAsync(D, (size_t)ldd * 8, dD, (size_t)w * 8, (size_t)w * 8, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * 8, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
