// panel.cu -- K1/K2: the latency-bound pieces of the blocked QL.
//
//  strip_kernel      bottom-up Householder QL (LAPACK ?geql2 + ?larfg conventions, the arithmetic the
//                    reference reaches through LAPACK.geqlf!, src/ql.jl:72; unblocked spec
//                    src/ql.jl:56-68) of a tall strip of <= SB columns.  One thread-block CLUSTER owns
//                    the strip: every thread keeps RPT rows x SB columns in registers, so the strip is
//                    read from and written to global memory exactly once.  Per column ONE fused
//                    reduction (norm of the pivot column + its dot products with every other column of
//                    the strip + the row of the pivot) goes warp shuffles -> shared memory ->
//                    distributed shared memory of every CTA in the cluster -> one hardware cluster
//                    barrier.  All sums are taken in a fixed order (bitwise reproducible).
//                    Also emits the strip's compact-WY T (lower triangular, ?larft 'B','C') and the
//                    "clean" reflector block Vw (explicit unit diagonal, zeros below) used by the GEMMs.
//  reduce_mult_t     W2 = (sum of split-K partials) * T_strip for the in-panel strip update.
//  larft_kernel      T (ib x ib lower) from the Gram matrix G = V'V and tau (?larft backward/columnwise
//                    recurrence), in place in shared memory.
//  extract_v_kernel  Vw from packed factors (for ormql).
#include "common.cuh"
#include "internal.h"

namespace qlb {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f64(const void *local_smem_ptr, uint32_t cta, double v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem_ptr)), "r"(cta));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}

struct StripParams {
    double *A;          // first active column of the strip, row 0
    int64_t lda;
    int ps;             // rows of the strip (pivot of the last column is row ps-1)
    int w;              // active columns (<= SB)
    double *tau;        // w entries
    double *Vw;         // clean reflectors out: column 0 of this strip inside the panel's Vw
    int64_t ldv;
    int p;              // rows of the panel (Vw rows [ps, p) are zero-filled)
    double *T;          // w x w lower-triangular T of the strip, leading dimension ldt
    int ldt;
    int rpc;            // rows per CTA (multiple of THREADS)
};

constexpr int STRIP_THREADS = 256;
constexpr int STRIP_MAXC = 16;      // largest cluster

template <int SB, int RPT>
__global__ void __launch_bounds__(STRIP_THREADS, 1) strip_kernel(const StripParams p) {
    constexpr int NW = STRIP_THREADS / 32;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ double wpart[NW][SB];                 // per-warp partial dots
    __shared__ double rowv_local[SB];                // row of the pivot (owner CTA only)
    __shared__ double red[2][STRIP_MAXC][SB];        // partial dots of every CTA of the cluster (DSMEM target)
    __shared__ double rowv[2][SB];                   // pivot row, broadcast by the owner CTA (DSMEM target)
    __shared__ double tot[SB];                       // cluster-wide totals
    __shared__ double Gs[SB][SB + 1];                // v_j . v_c for the T recurrence (CTA 0)
    __shared__ double Ts[SB][SB + 1];
    __shared__ double taus[SB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta = cluster_ctarank(), nc = cluster_nctarank();
    const int off = SB - p.w;                        // local columns [off, SB) are active
    constexpr int NOROW = 0x3fffffff;                // rows this thread does not own compare false everywhere
    int rr[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = (int)cta * p.rpc + tid + i * STRIP_THREADS;
        rr[i] = (r < p.ps && r < (int)(cta + 1) * p.rpc) ? r : NOROW;
    }

    double a[RPT][SB];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            const bool ok = (rr[i] != NOROW) && (j >= off);
            a[i][j] = ok ? p.A[rr[i] + (int64_t)(j - off) * p.lda] : 0.0;
        }
    }
    if (tid < SB) taus[tid] = 0.0;
    cluster_sync_all();          // every CTA of the cluster is resident before any remote store

    int par = 0;
    for (int c = SB - 1; c >= off; --c, par ^= 1) {
        const int mu = p.ps - (SB - c);              // pivot row of local column c
        // ---- 1. this thread's part of d_j = sum_{r < mu} x_r a_j[r], and the pivot row -------------
        double x[RPT], d[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) d[j] = 0.0;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int r = rr[i];
            double xi = 0.0;
#pragma unroll
            for (int j = 0; j < SB; ++j) xi = (j == c) ? a[i][j] : xi;
            if (r == mu) {
#pragma unroll
                for (int j = 0; j < SB; ++j) rowv_local[j] = a[i][j];
            }
            x[i] = (r < mu) ? xi : 0.0;
#pragma unroll
            for (int j = 0; j < SB; ++j) d[j] = fma(x[i], a[i][j], d[j]);
        }
        // ---- 2. warp reduction of SB values (recursive halving, fixed order) ------------------------
        // after the loop lane L holds the warp total of value index vidx(L)
        {
            int n = SB;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                if (n > 1) {
                    const bool up = (lane & o) != 0;
                    n >>= 1;
#pragma unroll
                    for (int q = 0; q < SB / 2; ++q) {
                        if (q < n) {
                            const double keep = up ? d[q + n] : d[q];
                            const double send = up ? d[q] : d[q + n];
                            d[q] = keep + __shfl_xor_sync(FULL, send, o);
                        }
                    }
                } else {
                    d[0] += __shfl_xor_sync(FULL, d[0], o);
                }
            }
        }
        // index of the value this lane ended up with: every halving round added (up ? n_round : 0)
        int vidx;
        {
            int n = SB, v = 0;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                if (n > 1) {
                    n >>= 1;
                    if (lane & o) v += n;
                }
            }
            vidx = v;
        }
        constexpr int LANES_PER_VAL = 32 / SB;       // lanes holding the same total (SB <= 32)
        if ((lane & (LANES_PER_VAL - 1)) == 0) wpart[warp][vidx] = d[0];
        __syncthreads();
        // ---- 3. CTA partials -> every CTA of the cluster (distributed shared memory) ----------------
        const int owner = mu / p.rpc;
        if (tid < SB * (int)nc) {
            const int j = tid % SB;
            const uint32_t dst = tid / SB;
            double v = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < NW; ++w8) v += wpart[w8][j];
            st_cluster_f64(&red[par][cta][j], dst, v);
            if ((int)cta == owner) st_cluster_f64(&rowv[par][j], dst, rowv_local[j]);
        }
        cluster_sync_all();
        // ---- 4. totals --------------------------------------------------------------------------------
        if (tid < SB) {
            double v = 0.0;
            for (uint32_t k = 0; k < nc; ++k) v += red[par][k][tid];
            tot[tid] = v;
        }
        __syncthreads();
        const double ssq = tot[c], alpha = rowv[par][c];
        // ---- 5. ?larfg scalars -------------------------------------------------------------------------
        double beta = alpha, tauc = 0.0, scale = 0.0;
        if (ssq != 0.0) {
            const double nrm = sqrt(fma(alpha, alpha, ssq));
            beta = -copysign(nrm, alpha);
            tauc = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        // ---- 6. rank-1 update of the columns left of c; column c becomes v (tail) and beta ------------
        double vi[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) vi[i] = (rr[i] < mu) ? x[i] * scale : ((rr[i] == mu) ? 1.0 : 0.0);
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            const double coef = (j < c) ? -tauc * fma(tot[j], scale, rowv[par][j]) : 0.0;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const double upd = fma(coef, vi[i], a[i][j]);
                const double colc = (rr[i] < mu) ? vi[i] : ((rr[i] == mu) ? beta : a[i][j]);
                a[i][j] = (j == c) ? colc : upd;
            }
        }
        if (cta == 0) {
            if (tid < SB && tid > c) Gs[tid][c] = fma(scale, tot[tid], rowv[par][tid]);      // v_tid . v_c
            if (tid == 0) taus[c] = tauc;
        }
        // tot / wpart / rowv_local are rewritten only after the next __syncthreads / cluster barrier
    }

    // ---- write back the factors and the clean reflectors ----------------------------------------------
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rr[i];
        if (r != NOROW) {
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                if (j >= off) {
                    const int mu = p.ps - (SB - j);
                    p.A[r + (int64_t)(j - off) * p.lda] = a[i][j];
                    p.Vw[r + (int64_t)(j - off) * p.ldv] = (r < mu) ? a[i][j] : ((r == mu) ? 1.0 : 0.0);
                }
            }
        }
    }
    if (cta == 0) {
        for (int j = 0; j < p.w; ++j)
            for (int r = p.ps + tid; r < p.p; r += STRIP_THREADS) p.Vw[r + (int64_t)j * p.ldv] = 0.0;
        __syncthreads();
        // T of the strip: backward/columnwise ?larft recurrence on the SB x SB Gram entries (warp 0)
        if (warp == 0) {
            for (int j = SB - 1; j >= off; --j) {
                const double tj = taus[j];
                double s = 0.0;
                if (lane < SB && lane > j)
                    for (int l = j + 1; l <= lane; ++l) s = fma(Ts[lane][l], Gs[l][j], s);
                __syncwarp();
                if (lane < SB) Ts[lane][j] = (lane > j) ? -tj * s : ((lane == j) ? tj : 0.0);
                __syncwarp();
            }
            for (int idx = lane; idx < p.w * p.w; idx += 32) {
                const int i = idx % p.w, j = idx / p.w;
                p.T[i + j * p.ldt] = Ts[i + off][j + off];
            }
            if (lane < p.w) p.tau[lane] = taus[lane + off];
        }
    }
}

// W2 (rows x w, ldw) = (sum_{s<S} part[s*stride + i + l*ldp]) * T (w x w, lower, ldt), w <= 16.
// blockDim = (16, 16): threadIdx.x = local row (coalesced loads), threadIdx.y = column l.
__global__ void reduce_mult_t_kernel(const double *__restrict__ part, int S, int64_t stride, int rows, int w, int64_t ldp,
                                     const double *__restrict__ T, int ldt, double *__restrict__ W2, int64_t ldw) {
    __shared__ double sums[16][17];
    __shared__ double Tsm[16][17];
    const int ri = threadIdx.x, l = threadIdx.y;
    const int i = blockIdx.x * 16 + ri;
    Tsm[ri][l] = (ri < w && l < w && ri >= l) ? T[ri + l * ldt] : 0.0;      // Tsm[row][col]
    double s = 0.0;
    if (i < rows && l < w) {
        const double *src = part + i + (int64_t)l * ldp;
        for (int z = 0; z < S; ++z) s += src[(int64_t)z * stride];
    }
    sums[ri][l] = s;
    __syncthreads();
    if (i < rows && l < w) {
        double o = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) o = fma(sums[ri][q], Tsm[q][l], o);
        W2[i + (int64_t)l * ldw] = o;
    }
}

// T (ib x ib, lower triangular, zero above) from G = V'V (ib x ib, only the strict lower part is read)
// and tau; T(j,j) = tau_j, T(j+1:,j) = -tau_j * T(j+1:,j+1:) * G(j+1:,j), j = ib-1 ... 0.
// One CTA; the recurrence runs in place in shared memory (pitch ib).
__global__ void __launch_bounds__(256, 1) larft_kernel(const double *__restrict__ G, int64_t ldg, const double *__restrict__ tau,
                                                      int ib, double *__restrict__ T, int64_t ldt) {
    extern __shared__ __align__(16) double sm[];
    double *W = sm;                     // ib x ib, column-major pitch ib
    double *ts = sm + (size_t)ib * ib;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < ib * ib; idx += blockDim.x) {
        const int i = idx % ib, j = idx / ib;
        W[idx] = (i > j) ? G[i + (int64_t)j * ldg] : 0.0;
    }
    if (tid < ib) ts[tid] = tau[tid];
    __syncthreads();
    // two threads per row: thread (i, h) sums l = j+1+h, j+3+h, ... ; combined through a shuffle
    const int i = tid >> 1, hh = tid & 1;
    for (int j = ib - 1; j >= 0; --j) {
        double s0 = 0.0, s1 = 0.0;
        if (i < ib && i > j) {
            int l = j + 1 + hh;
            for (; l + 2 <= i; l += 4) {
                s0 = fma(W[l * ib + i], W[j * ib + l], s0);
                s1 = fma(W[(l + 2) * ib + i], W[j * ib + l + 2], s1);
            }
            if (l <= i) s0 = fma(W[l * ib + i], W[j * ib + l], s0);
        }
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        __syncthreads();
        if (hh == 0 && i < ib) {
            if (i > j) W[j * ib + i] = -ts[j] * s;
            else if (i == j) W[j * ib + i] = ts[j];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < ib * ib; idx += blockDim.x) {
        const int r = idx % ib, c = idx / ib;
        T[r + (int64_t)c * ldt] = (r >= c) ? W[idx] : 0.0;
    }
}

// Vw (p x ib, ldv) from the packed factors F (column j has its pivot at row mu0 + j).
__global__ void extract_v_kernel(const double *__restrict__ F, int64_t ldf, int p, int ib, int mu0, double *__restrict__ Vw,
                                 int64_t ldv) {
    const int j = blockIdx.y;
    const int mu = mu0 + j;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p; r += gridDim.x * blockDim.x)
        Vw[r + (int64_t)j * ldv] = (r < mu) ? F[r + (int64_t)j * ldf] : ((r == mu) ? 1.0 : 0.0);
}

template <int SB, int RPT>
static int launch_strip_t(qlb200_ctx *h, cudaStream_t st, const StripParams &p, int nc) {
    auto kern = strip_kernel<SB, RPT>;
    if (!h->strip_attr_done[RPT]) {
        QLB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        h->strip_attr_done[RPT] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc, 1, 1);
    cfg.blockDim = dim3(STRIP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    QLB_CUDA(h, cudaLaunchKernelEx(&cfg, kern, p));
    h->launches++;
    return 0;
}

}  // namespace qlb

// Largest strip (rows) the cluster kernel can hold: 16 CTAs x 256 threads x 4 rows.
int qlb_strip_max_rows() { return qlb::STRIP_MAXC * qlb::STRIP_THREADS * 4; }

int qlb_strip_factor(qlb200_ctx *h, cudaStream_t st, double *A, int64_t lda, int ps, int w, double *tau, double *Vw,
                     int64_t ldv, int p, double *T, int ldt) {
    using namespace qlb;
    if (w < 1 || w > 16) return qlb_fail(h, -2, "strip width must be in [1,16]%s");
    if (ps > qlb_strip_max_rows()) return qlb_fail(h, -2, "strip taller than the cluster kernel supports%s");
    // rows per thread, then CTAs per cluster (power of two <= 16)
    int rpt = (ps + STRIP_THREADS - 1) / STRIP_THREADS;
    rpt = rpt <= 1 ? 1 : (rpt <= 2 ? 2 : 4);
    int nc = (ps + STRIP_THREADS * rpt - 1) / (STRIP_THREADS * rpt);
    int ncp = 1;
    while (ncp < nc) ncp <<= 1;
    nc = ncp;
    StripParams p_{A, lda, ps, w, tau, Vw, ldv, p, T, ldt, STRIP_THREADS * rpt};
    // spread the rows evenly over the CTAs (rpc multiple of the block size)
    int rpc = ((ps + nc - 1) / nc + STRIP_THREADS - 1) / STRIP_THREADS * STRIP_THREADS;
    if (rpc > STRIP_THREADS * rpt) rpc = STRIP_THREADS * rpt;
    p_.rpc = rpc;
    switch (rpt) {
        case 1: return launch_strip_t<16, 1>(h, st, p_, nc);
        case 2: return launch_strip_t<16, 2>(h, st, p_, nc);
        default: return launch_strip_t<16, 4>(h, st, p_, nc);
    }
}

int qlb_reduce_mult_t(qlb200_ctx *h, cudaStream_t st, const double *part, int S, int64_t stride, int rows, int w,
                      int64_t ldp, const double *T, int ldt, double *W2, int64_t ldw) {
    if (rows <= 0 || w <= 0) return 0;
    dim3 block(16, 16);
    qlb::reduce_mult_t_kernel<<<(rows + 15) / 16, block, 0, st>>>(part, S, stride, rows, w, ldp, T, ldt, W2, ldw);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

int qlb_larft(qlb200_ctx *h, cudaStream_t st, const double *G, int64_t ldg, const double *tau, int ib, double *T,
              int64_t ldt) {
    if (ib <= 0) return 0;
    if (ib > 128) return qlb_fail(h, -2, "larft: block wider than 128%s");
    const size_t smem = ((size_t)ib * ib + ib) * sizeof(double);
    if (!h->larft_attr_done) {
        QLB_CUDA(h, cudaFuncSetAttribute(qlb::larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 * 128 + 128) * 8));
        h->larft_attr_done = true;
    }
    qlb::larft_kernel<<<1, 256, smem, st>>>(G, ldg, tau, ib, T, ldt);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

int qlb_extract_v(qlb200_ctx *h, cudaStream_t st, const double *F, int64_t ldf, int p, int ib, int mu0, double *Vw,
                  int64_t ldv) {
    if (p <= 0 || ib <= 0) return 0;
    dim3 grid((p + 255) / 256, ib);
    if (grid.x > 64) grid.x = 64;
    qlb::extract_v_kernel<<<grid, 256, 0, st>>>(F, ldf, p, ib, mu0, Vw, ldv);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
