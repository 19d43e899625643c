// ul.cu -- A13: the reverse-order partially pivoted LU of the reference, A[p,:] = U*L with U unit upper and
// L lower triangular (generic_ulfact!, reference src/ul.jl:148-196), and ldiv!(F::UL, B) (src/ul.jl:363-393).
//
// Blocked right-looking elimination that starts at the bottom-right corner, strips of SB = 16 columns:
//   ul_panel_kernel   columns [c0,c1) x rows [0,c1): per column k = c1-1 ... c0 the pivot search over rows 0..k
//                     (first row attaining the maximum, exactly the reference's strict '>' scan, src/ul.jl:158-166),
//                     the row interchange inside the strip, the scaling of the column above the pivot
//                     (src/ul.jl:177-180) and the rank-1 update of the strip's remaining columns (:185-189);
//   ul_laswp_kernel   the same interchanges on every column outside the strip (the reference swaps whole rows, :170-175);
//   ul_trsm_row_kernel  rows [c0,c1) of the columns left of the strip: L21 = U22^{-1} A21 (the updates :185-189 that
//                     the strip's own columns apply to those rows, collected as one unit-upper back substitution);
//   DMMA GEMM         A11 -= U12 * L21 (gemm.cu).
// ipiv is 1-based like the reference's Vector{BlasInt}; info = the largest k with a zero pivot (the first one the
// descending loop meets, src/ul.jl:181-183), 0 if none.
// Two-level blocking: strips of 16 update only their 128-column outer block; everything left of the block gets one
// rank-128 DMMA update per block.  Round-1 status: the strip kernel is a single CTA (latency-bound) -- next: a cluster.
#include "common.cuh"
#include "internal.h"

namespace qlb {

constexpr int UL_SB = 16;
constexpr int UL_THREADS = 1024;

struct ArgMax {
    double v;
    int i;
};
__device__ __forceinline__ ArgMax argmax_combine(ArgMax a, ArgMax b) {
    // larger value wins; on ties the lower row index (the reference keeps the first maximum)
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

__global__ void __launch_bounds__(UL_THREADS) ul_panel_kernel(double *__restrict__ A, int64_t lda, int c0, int c1, int pivot,
                                                              int64_t *__restrict__ ipiv, int32_t *__restrict__ info) {
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    __shared__ double rowk[UL_SB];
    __shared__ int s_kp;
    __shared__ double s_piv;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = c1 - 1; k >= c0; --k) {
        double *colk = A + (int64_t)k * lda;
        // ---- pivot search over rows 0..k of column k ----
        if (pivot) {
            ArgMax best{0.0, 0x7fffffff};
            for (int r = tid; r <= k; r += UL_THREADS) {
                const double v = fabs(colk[r]);
                if (v > best.v) best = ArgMax{v, r};          // ascending r inside a thread: first maximum kept
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
                best = argmax_combine(best, other);
            }
            if (lane == 0) { red_v[warp] = best.v; red_i[warp] = best.i; }
            __syncthreads();
            if (warp == 0) {
                ArgMax b{red_v[lane], red_i[lane]};
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    ArgMax other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
                    b = argmax_combine(b, other);
                }
                if (lane == 0) s_kp = (b.i == 0x7fffffff) ? k : b.i;       // all zero (or NaN): kp = k
            }
        } else if (tid == 0) {
            s_kp = k;
        }
        __syncthreads();
        const int kp = s_kp;
        if (tid == 0) {
            ipiv[k] = (int64_t)kp + 1;
            s_piv = colk[kp];
        }
        __syncthreads();
        const double piv = s_piv;
        if (piv != 0.0) {
            // ---- interchange rows k and kp inside the strip; the new row k feeds the update ----
            if (tid < c1 - c0) {
                double *col = A + (int64_t)(c0 + tid) * lda;
                const double a = col[k], b = col[kp];
                col[k] = b;
                col[kp] = a;
                rowk[tid] = b;
            }
        } else {
            if (tid == 0 && *info == 0) *info = k + 1;
            if (tid < c1 - c0) rowk[tid] = A[(int64_t)(c0 + tid) * lda + k];
        }
        __syncthreads();
        // ---- scale the column above the pivot, rank-1 update of the strip's columns left of k ----
        const double pinv = (piv != 0.0) ? 1.0 / piv : 1.0;
        const int nj = k - c0;
        for (int r = tid; r < k; r += UL_THREADS) {
            double m = colk[r];
            if (piv != 0.0) {
                m *= pinv;
                colk[r] = m;
            }
            for (int j = 0; j < nj; ++j) {
                double *p = A + (int64_t)(c0 + j) * lda + r;
                *p = fma(-m, rowk[j], *p);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster version of the strip kernel: the strip (rows [0,c1) x <= 16 columns) lives in the REGISTERS of one
// thread-block cluster (<= 16 CTAs x 512 threads x 2 rows), so it is read and written once; per column two
// hardware cluster barriers: (1) every CTA's pivot candidate has reached every CTA through distributed shared
// memory, (2) the two rows to interchange have been published by their owners.  Same arithmetic and the same
// first-maximum pivot rule as ul_panel_kernel (reference src/ul.jl:158-189).
// ---------------------------------------------------------------------------------------------
constexpr int ULC_THREADS = 512;
constexpr int ULC_RPT = 2;
constexpr int ULC_MAXC = 16;

__device__ __forceinline__ uint32_t ulc_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t ulc_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void ulc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ulc_map(const void *local_smem_ptr, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem_ptr)), "r"(cta));
    return ra;
}
__device__ __forceinline__ void ulc_st_f64(uint32_t addr, double v) { asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }
__device__ __forceinline__ void ulc_st_s32(uint32_t addr, int v) { asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ double ulc_ld_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

struct UlcSmem {
    double red_v[ULC_THREADS / 32];
    int red_i[ULC_THREADS / 32];
    double cand_v[2][ULC_MAXC];          // every CTA's candidate, written by its owner into all CTAs (double-buffered by step parity)
    int cand_i[2][ULC_MAXC];
    double pub[2][2][UL_SB];             // [parity][0: row kp, 1: row k][column]: rows published by their owner threads
    double rowk_new[UL_SB], rowk_old[UL_SB];
};

__global__ void __launch_bounds__(ULC_THREADS, 1) ul_strip_cluster_kernel(double *__restrict__ A, int64_t lda, int c0, int c1, int pivot,
                                                                          int64_t *__restrict__ ipiv, int32_t *__restrict__ info) {
    __shared__ UlcSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = (int)ulc_ctarank(), nc = (int)ulc_nctarank();
    constexpr int RPC = ULC_THREADS * ULC_RPT;           // rows per CTA
    const int w = c1 - c0, off = UL_SB - w;              // local column lc = off + (column - c0); the last column sits at 15
    int rr[ULC_RPT];
    double a[ULC_RPT][UL_SB];
#pragma unroll
    for (int i = 0; i < ULC_RPT; ++i) {
        rr[i] = cta * RPC + tid + i * ULC_THREADS;
#pragma unroll
        for (int q = 0; q < UL_SB; ++q) a[i][q] = (rr[i] < c1 && q >= off) ? A[(int64_t)(c0 + q - off) * lda + rr[i]] : 0.0;
    }
    ulc_cluster_sync();            // every CTA of the cluster is resident before the first remote store
#pragma unroll
    for (int s = 0; s < UL_SB; ++s) {
        if (s < w) {
            constexpr int dummy = 0;
            (void)dummy;
            const int lc = UL_SB - 1 - s;          // compile-time register column of the pivot column
            const int k = c1 - 1 - s;              // its global index = pivot row
            const int par = s & 1;
            // ---- 1. pivot search: first row attaining max |a[r][k]|, r <= k ----
            int kp = k;
            if (pivot) {
                ArgMax best{0.0, 0x7fffffff};
#pragma unroll
                for (int i = 0; i < ULC_RPT; ++i) {
                    const double v = fabs(a[i][lc]);
                    if (rr[i] <= k && v > best.v) best = ArgMax{v, rr[i]};      // rr ascending in i
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    ArgMax other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
                    best = argmax_combine(best, other);
                }
                if (lane == 0) { sm.red_v[warp] = best.v; sm.red_i[warp] = best.i; }
                __syncthreads();
                if (warp == 0) {
                    ArgMax b{lane < ULC_THREADS / 32 ? sm.red_v[lane] : 0.0, lane < ULC_THREADS / 32 ? sm.red_i[lane] : 0x7fffffff};
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        ArgMax other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
                        b = argmax_combine(b, other);
                    }
                    if (lane < nc) {               // lane c delivers this CTA's candidate to CTA c
                        ulc_st_f64(ulc_map(&sm.cand_v[par][cta], (uint32_t)lane), b.v);
                        ulc_st_s32(ulc_map(&sm.cand_i[par][cta], (uint32_t)lane), b.i);
                    }
                }
                ulc_cluster_sync();
                ArgMax gbest{0.0, 0x7fffffff};
                for (int c = 0; c < nc; ++c) gbest = argmax_combine(gbest, ArgMax{sm.cand_v[par][c], sm.cand_i[par][c]});
                kp = (gbest.i == 0x7fffffff) ? k : gbest.i;
            }
            // ---- 2. the owners publish rows kp and k ----
#pragma unroll
            for (int i = 0; i < ULC_RPT; ++i) {
                if (rr[i] == kp) {
#pragma unroll
                    for (int q = 0; q < UL_SB; ++q) sm.pub[par][0][q] = a[i][q];
                }
                if (rr[i] == k) {
#pragma unroll
                    for (int q = 0; q < UL_SB; ++q) sm.pub[par][1][q] = a[i][q];
                }
            }
            ulc_cluster_sync();
            if (warp == 0) {
                const int src_cta = (lane < 16) ? kp / RPC : k / RPC;
                const double v = ulc_ld_f64(ulc_map(&sm.pub[par][lane < 16 ? 0 : 1][lane & 15], (uint32_t)src_cta));
                if (lane < 16) sm.rowk_new[lane] = v;
                else sm.rowk_old[lane & 15] = v;
            }
            __syncthreads();
            const double piv = sm.rowk_new[lc];
            if (cta == 0 && tid == 0) {
                ipiv[k] = (int64_t)kp + 1;
                if (piv == 0.0 && *info == 0) *info = k + 1;
            }
            // ---- 3. interchange (registers), scale the column above the pivot, rank-1 update to its left ----
            const bool doswap = (piv != 0.0) && (kp != k);
            const double pinv = (piv != 0.0) ? 1.0 / piv : 1.0;
#pragma unroll
            for (int i = 0; i < ULC_RPT; ++i) {
                if (doswap && rr[i] == k) {
#pragma unroll
                    for (int q = 0; q < UL_SB; ++q) a[i][q] = sm.rowk_new[q];
                }
                if (doswap && rr[i] == kp) {
#pragma unroll
                    for (int q = 0; q < UL_SB; ++q) a[i][q] = sm.rowk_old[q];
                }
                if (rr[i] < k) {
                    double m = a[i][lc];
                    if (piv != 0.0) {
                        m *= pinv;
                        a[i][lc] = m;
                    }
                    // the pivot row the reference uses for the update is row k AFTER the interchange (or the untouched
                    // row k when the pivot is zero and nothing is interchanged)
#pragma unroll
                    for (int q = 0; q < UL_SB; ++q)
                        if (q < lc) a[i][q] = fma(-m, (piv != 0.0) ? sm.rowk_new[q] : sm.rowk_old[q], a[i][q]);
                }
            }
            __syncthreads();       // rowk_new / rowk_old / red_* are rewritten in the next step
        }
    }
#pragma unroll
    for (int i = 0; i < ULC_RPT; ++i) {
        if (rr[i] < c1) {
#pragma unroll
            for (int q = 0; q < UL_SB; ++q)
                if (q >= off) A[(int64_t)(c0 + q - off) * lda + rr[i]] = a[i][q];
        }
    }
    ulc_cluster_sync();            // no CTA exits while its shared memory may still be read remotely
}

// Row interchanges of one strip (k = c1-1 ... c0, in that order) applied to the columns [j0, j1).
__global__ void ul_laswp_kernel(double *__restrict__ A, int64_t lda, int c0, int c1, const int64_t *__restrict__ ipiv, int j0,
                                int j1) {
    const int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= j1) return;
    double *col = A + (int64_t)j * lda;
    for (int k = c1 - 1; k >= c0; --k) {
        const int kp = (int)ipiv[k] - 1;
        if (kp != k) {
            const double a = col[k], b = col[kp];
            col[k] = b;
            col[kp] = a;
        }
    }
}

// X = U22^{-1} A21 in place: U22 = unit upper triangle of A[c0:c1, c0:c1], A21 = A[c0:c1, col0:col0+ncols).
__global__ void ul_trsm_row_kernel(double *__restrict__ A, int64_t lda, int c0, int c1, int col0, int ncols) {
    __shared__ double U[UL_SB][UL_SB + 1];
    const int w = c1 - c0;
    for (int idx = threadIdx.x; idx < w * w; idx += blockDim.x) {
        const int i = idx % w, j = idx / w;
        U[i][j] = A[(int64_t)(c0 + j) * lda + c0 + i];
    }
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    double *x = A + (int64_t)(col0 + j) * lda + c0;
    double v[UL_SB];
#pragma unroll
    for (int i = 0; i < UL_SB; ++i) v[i] = (i < w) ? x[i] : 0.0;
#pragma unroll
    for (int i = UL_SB - 1; i >= 0; --i) {
        if (i < w) {
#pragma unroll
            for (int q = i + 1; q < UL_SB; ++q)
                if (q < w) v[i] = fma(-U[i][q], v[q], v[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < UL_SB; ++i)
        if (i < w) x[i] = v[i];
}

// Rows of B permuted by ipiv in REVERSE order (_apply_ipiv_rows!, src/ul.jl:363-376): thread = column of B.
__global__ void ul_permute_rows_kernel(double *__restrict__ B, int64_t ldb, int n, int nrhs, const int64_t *__restrict__ ipiv) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nrhs) return;
    double *col = B + (int64_t)c * ldb;
    for (int i = n - 1; i >= 0; --i) {
        const int p = (int)ipiv[i] - 1;
        if (p != i) {
            const double a = col[i], b = col[p];
            col[i] = b;
            col[p] = a;
        }
    }
}

// Back substitution with one tb x tb UNIT upper-triangular diagonal block: X = U^{-1} B in place.
template <int TB>
__global__ void __launch_bounds__(TB) trsm_unit_upper_diag_kernel(const double *__restrict__ U, int64_t ldu, int tb, double *__restrict__ B,
                                                                 int64_t ldb, int nrhs) {
    extern __shared__ __align__(16) double sm[];
    double *Us = sm;                       // tb x tb, pitch TB+1
    __shared__ double xs[8];
    const int r = threadIdx.x;
    for (int idx = threadIdx.x; idx < tb * tb; idx += TB) {
        const int i = idx % tb, j = idx / tb;
        Us[j * (TB + 1) + i] = (i < j) ? U[i + (int64_t)j * ldu] : 0.0;
    }
    __syncthreads();
    for (int c0 = blockIdx.x * 8; c0 < nrhs; c0 += gridDim.x * 8) {
        const int nc = (nrhs - c0) < 8 ? (nrhs - c0) : 8;
        double b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] = (r < tb && q < nc) ? B[r + (int64_t)(c0 + q) * ldb] : 0.0;
        for (int c = tb - 1; c >= 0; --c) {
            if (r == c) {
#pragma unroll
                for (int q = 0; q < 8; ++q) xs[q] = b[q];
            }
            __syncthreads();
            if (r < c) {
                const double u = Us[c * (TB + 1) + r];
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] = fma(-u, xs[q], b[q]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (r < tb && q < nc) B[r + (int64_t)(c0 + q) * ldb] = b[q];
    }
}

__global__ void ul_set_info_kernel(int32_t *p, int32_t v) { *p = v; }

}  // namespace qlb

using namespace qlb;

// ipiv: device array of min(m,n) int64 (1-based); dinfo: device int32 (0 or the reference's `info`).
int qlb_dulfact_dev(qlb200_ctx *h, int64_t m64, int64_t n64, double *A, int64_t lda, int64_t *ipiv, int pivot, int32_t *dinfo) {
    const int m = (int)m64, n = (int)n64;
    const int k = m < n ? m : n;
    cudaStream_t st = h->stream;
    ul_set_info_kernel<<<1, 1, 0, st>>>(dinfo, 0);
    QLB_LAUNCH_CHECK(h);
    // The reference's loop touches rows/columns 1..k only (k = min(m,n) ... 1), except for the row
    // interchanges which run over all n columns (src/ul.jl:155-191).
    // Two levels: outer blocks of UL_NBO = 128 columns; inside a block the strips update only the block's own columns
    // (K = 16, small); the columns left of the block get ONE rank-128 DMMA update per block, after the block's rows of L
    // have been finished by a block back substitution with the block's unit upper triangle.
    constexpr int UL_NBO = 128;
    // One outer block [C0, C1): its strips from the right, the row interchanges applied at once only INSIDE the block's own
    // columns (nothing outside the block is read until the block is complete, so the interchanges of the other columns are
    // deferred and applied in one pass, see below), the in-block rows of L and the in-block rank-16 updates.
    auto factor_block = [&](cudaStream_t s, int C0, int C1) -> int {
        for (int c1 = C1; c1 > C0; c1 -= UL_SB) {
            const int c0 = (c1 - UL_SB > C0) ? c1 - UL_SB : C0;
            if (c1 <= ULC_MAXC * ULC_THREADS * ULC_RPT && h->opt_ul_cluster) {
                int nc = (c1 + ULC_THREADS * ULC_RPT - 1) / (ULC_THREADS * ULC_RPT), ncp = 1;
                while (ncp < nc) ncp <<= 1;
                if (!h->ul_cluster_attr_done) {      // per handle (= per device), like every other attribute
                    QLB_CUDA(h, cudaFuncSetAttribute(ul_strip_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
                    h->ul_cluster_attr_done = true;
                }
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(ncp, 1, 1);
                cfg.blockDim = dim3(ULC_THREADS, 1, 1);
                cfg.dynamicSmemBytes = 0;
                cfg.stream = s;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = ncp;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                QLB_CUDA(h, cudaLaunchKernelEx(&cfg, ul_strip_cluster_kernel, A, lda, c0, c1, pivot, ipiv, dinfo));
                h->launches++;
            } else {
                ul_panel_kernel<<<1, UL_THREADS, 0, s>>>(A, lda, c0, c1, pivot, ipiv, dinfo);
                QLB_LAUNCH_CHECK(h);
            }
            if (pivot) {
                if (c0 > C0) {
                    ul_laswp_kernel<<<(c0 - C0 + 255) / 256, 256, 0, s>>>(A, lda, c0, c1, ipiv, C0, c0);
                    QLB_LAUNCH_CHECK(h);
                }
                if (C1 > c1) {
                    ul_laswp_kernel<<<(C1 - c1 + 255) / 256, 256, 0, s>>>(A, lda, c0, c1, ipiv, c1, C1);
                    QLB_LAUNCH_CHECK(h);
                }
            }
            if (c0 > C0) {
                // inside the block: rows [c0,c1) of the block's columns left of the strip, then A11_block -= U12 * L21
                ul_trsm_row_kernel<<<(c0 - C0 + 127) / 128, 128, 0, s>>>(A, lda, c0, c1, C0, c0 - C0);
                QLB_LAUNCH_CHECK(h);
                int rc = qlb_gemm(h, s, -1, false, true, c0, c0 - C0, c1 - c0, -1.0, A + (int64_t)c0 * lda, lda,
                                  A + (int64_t)C0 * lda + c0, lda, 1.0, A + (int64_t)C0 * lda, lda, 1, 0);
                if (rc) return rc;
            }
        }
        return 0;
    };
    // The block's interchanges on the columns outside the block (left of it: still to be updated; right of it: finished)
    auto deferred_swaps = [&](cudaStream_t s, int C0, int C1) -> int {
        if (!pivot) return 0;
        if (C0 > 0) {
            ul_laswp_kernel<<<(C0 + 255) / 256, 256, 0, s>>>(A, lda, C0, C1, ipiv, 0, C0);
            QLB_LAUNCH_CHECK(h);
        }
        if (n > C1) {
            ul_laswp_kernel<<<(n - C1 + 255) / 256, 256, 0, s>>>(A, lda, C0, C1, ipiv, C1, n);
            QLB_LAUNCH_CHECK(h);
        }
        return 0;
    };
    // rows [C0,C1) of the columns left of the block: X = U_block^{-1} A21, strip by strip from the bottom
    auto block_rows = [&](cudaStream_t s, int C0, int C1) -> int {
        for (int c1 = C1; c1 > C0; c1 -= UL_SB) {
            const int c0 = (c1 - UL_SB > C0) ? c1 - UL_SB : C0;
            if (c1 < C1) {      // A21[c0:c1, :] -= U[c0:c1, c1:C1] * X[c1:C1, :]
                int rc = qlb_gemm(h, s, -1, false, true, c1 - c0, C0, C1 - c1, -1.0, A + (int64_t)c1 * lda + c0, lda, A + c1, lda,
                                  1.0, A + c0, lda, 1, 0);
                if (rc) return rc;
            }
            ul_trsm_row_kernel<<<(C0 + 127) / 128, 128, 0, s>>>(A, lda, c0, c1, 0, C0);
            QLB_LAUNCH_CHECK(h);
        }
        return 0;
    };
    // A11[0:C0, j0:j1) -= U12 (C0 x nbo, M-contiguous) * L21[:, j0:j1) (nbo x cols, K-contiguous)
    auto trailing = [&](cudaStream_t s, int C0, int C1, int j0, int j1) -> int {
        if (j1 <= j0) return 0;
        return qlb_gemm(h, s, -1, false, true, C0, j1 - j0, C1 - C0, -1.0, A + (int64_t)C0 * lda, lda, A + (int64_t)j0 * lda + C0, lda, 1.0,
                        A + (int64_t)j0 * lda, lda, 1, 0);
    };
    // Look-ahead on the SM partitions (as in ql!, blocked.cu): block b+1 is factored on the 16-SM panel partition while the
    // other partition applies block b to the columns left of block b+1 -- the 16 384 latency-bound column steps of the strips
    // (about half of the run time at n = 16384) disappear behind the rank-128 updates.
    const bool la = h->gc_ok && h->opt_lookahead != 0 && k >= 2048;
    if (!la) {
        for (int C1 = k; C1 > 0; C1 -= UL_NBO) {
            const int C0 = (C1 - UL_NBO > 0) ? C1 - UL_NBO : 0;
            int rc = factor_block(st, C0, C1);
            if (!rc) rc = deferred_swaps(st, C0, C1);
            if (!rc && C0 > 0) rc = block_rows(st, C0, C1);
            if (!rc && C0 > 0) rc = trailing(st, C0, C1, 0, C0);
            if (rc) return rc;
        }
        return 0;
    }
    cudaStream_t G = h->gc_gemm, P = h->gc_panel;
    cudaEvent_t evP = h->ev_a, evN = h->ev_b;
    QLB_CUDA(h, cudaEventRecord(h->ev_gc[0], st));
    QLB_CUDA(h, cudaStreamWaitEvent(G, h->ev_gc[0], 0));
    QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_gc[0], 0));
    {
        const int C1 = k, C0 = (C1 - UL_NBO > 0) ? C1 - UL_NBO : 0;
        int rc = factor_block(P, C0, C1);
        if (rc) return rc;
        QLB_CUDA(h, cudaEventRecord(evP, P));
    }
    for (int C1 = k; C1 > 0; C1 -= UL_NBO) {
        const int C0 = (C1 - UL_NBO > 0) ? C1 - UL_NBO : 0;
        QLB_CUDA(h, cudaStreamWaitEvent(G, evP, 0));            // block [C0, C1) is factored
        int rc = deferred_swaps(G, C0, C1);
        if (rc) return rc;
        if (C0 > 0) {
            rc = block_rows(G, C0, C1);
            if (rc) return rc;
            const int D0 = (C0 - UL_NBO > 0) ? C0 - UL_NBO : 0;      // next block = columns [D0, C0)
            rc = trailing(G, C0, C1, D0, C0);
            if (rc) return rc;
            QLB_CUDA(h, cudaEventRecord(evN, G));
            QLB_CUDA(h, cudaStreamWaitEvent(P, evN, 0));
            rc = factor_block(P, D0, C0);
            if (rc) return rc;
            QLB_CUDA(h, cudaEventRecord(evP, P));
            rc = trailing(G, C0, C1, 0, D0);
            if (rc) return rc;
        }
    }
    QLB_CUDA(h, cudaEventRecord(h->ev_gc[1], G));
    QLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_gc[1], 0));
    return 0;
}

// ldiv!(F::UL, B), square: permute rows (reverse ipiv order), unit-upper back substitution, lower forward substitution.
int qlb_dulsolve_dev(qlb200_ctx *h, int64_t n64, int64_t nrhs64, const double *A, int64_t lda, const int64_t *ipiv, double *B,
                     int64_t ldb, int32_t *dinfo) {
    const int n = (int)n64, nrhs = (int)nrhs64;
    if (n == 0 || nrhs == 0) return 0;
    cudaStream_t st = h->stream;
    ul_permute_rows_kernel<<<(nrhs + 127) / 128, 128, 0, st>>>(B, ldb, n, nrhs, ipiv);
    QLB_LAUNCH_CHECK(h);
    constexpr int TB = 128;
    const size_t smem = (size_t)TB * (TB + 1) * sizeof(double);
    if (!h->ul_trsm_attr_done) {
        QLB_CUDA(h, cudaFuncSetAttribute(trsm_unit_upper_diag_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->ul_trsm_attr_done = true;
    }
    const int nblk = (n + TB - 1) / TB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int i = b * TB, tb = (n - i) < TB ? (n - i) : TB;
        trsm_unit_upper_diag_kernel<TB><<<trsm_grid(h, nrhs), TB, smem, st>>>(A + i + (int64_t)i * lda, lda, tb, B + i, ldb, nrhs);
        QLB_LAUNCH_CHECK(h);
        if (i > 0) {      // B[0:i] -= U[0:i, i:i+tb] * X[i:i+tb]
            int rc = qlb_gemm(h, st, nrhs <= 16 ? 2 : -1, false, true, i, nrhs, tb, -1.0, A + (int64_t)i * lda, lda, B + i, ldb, 1.0, B, ldb,
                              1, 0);
            if (rc) return rc;
        }
    }
    return qlb_trsm_lower_dev(h, n, nrhs, A, lda, B, ldb, dinfo);
}
