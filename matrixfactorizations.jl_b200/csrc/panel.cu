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
// Asynchronous 8-byte store into the shared memory of CTA `cta` of the cluster; completion is
// signalled as 8 transaction bytes on that CTA's mbarrier (st.async + mbarrier::complete_tx).
__device__ __forceinline__ void st_async_f64(const void *local_smem_ptr, const uint64_t *local_mbar, uint32_t cta, double v) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem_ptr)), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_mbar)), "r"(cta));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(ra), "d"(v), "r"(rb)
                 : "memory");
}
// 8-byte load from the shared memory of CTA `cta` of the cluster (DSMEM).
__device__ __forceinline__ double ld_dsmem_f64(const void *local_smem_ptr, uint32_t cta) {
    uint32_t ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem_ptr)), "r"(cta));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
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

// Shared-memory working set of one strip CTA.
template <int SB>
struct StripSmem {
    static constexpr int NW = STRIP_THREADS / 32;
    double wpart[2][NW][SB];              // per-warp partial dots (double-buffered by step parity)
    double rowv_local[2][SB];             // row of the pivot (owner CTA only)
    double red[2][STRIP_MAXC][SB];        // partial dots of every CTA of the cluster (st.async target)
    double rowv[2][SB];                   // pivot row, sent by the owner CTA (st.async target)
    double totw[NW][SB];                  // cluster-wide totals, one private copy per warp
    uint64_t mbar[2];                     // transaction barriers guarding red[par]/rowv[par]
    double sl_w[NW][SB];                  // ?larfg slow path (scaled norm, safmin loop): per-warp partials,
    double sl_c[SB];                      //   this CTA's partials (read by every CTA of the cluster through DSMEM),
    double sl_tot[SB];                    //   cluster-wide results
    double Gs[SB][SB + 1];                // v_j . v_c for the T recurrence (CTA 0)
    double Ts[SB][SB + 1];
    double taus[SB];
};

// Slow-path all-reduce of K <= SB values per thread over the whole cluster (sum, or max when MAX): warp shuffles ->
// shared memory -> hardware cluster barrier -> DSMEM reads in a fixed order (bitwise identical in every CTA).  Results
// in sm.sl_tot[0..K).  Must be called by every thread of every CTA of the cluster.  Only the rare ?larfg special cases
// come here (squares that under/overflow, exact zero tails), so two cluster barriers per call are acceptable.
template <int SB, int K, bool MAX>
__device__ __noinline__ void cluster_allreduce_slow(const double (&v)[K], StripSmem<SB> &sm, const int tid, const uint32_t nc) {
    constexpr int NW = STRIP_THREADS / 32;
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double t = v[k];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const double u = __shfl_xor_sync(FULL, t, o);
            t = MAX ? fmax(t, u) : t + u;
        }
        if (lane == 0) sm.sl_w[warp][k] = t;
    }
    __syncthreads();
    if (tid < K) {
        double t = sm.sl_w[0][tid];
#pragma unroll
        for (int w8 = 1; w8 < NW; ++w8) t = MAX ? fmax(t, sm.sl_w[w8][tid]) : t + sm.sl_w[w8][tid];
        sm.sl_c[tid] = t;
    }
    cluster_sync_all();
    if (tid < K) {
        double t = ld_dsmem_f64(&sm.sl_c[tid], 0);
        for (uint32_t c = 1; c < nc; ++c) {
            const double u = ld_dsmem_f64(&sm.sl_c[tid], c);
            t = MAX ? fmax(t, u) : t + u;
        }
        sm.sl_tot[tid] = t;
    }
    cluster_sync_all();          // every CTA has read sl_c (it may be rewritten now); sl_tot is visible to the whole CTA
}

__device__ __forceinline__ double dlapy2_dev(double x, double y) {
    const double xa = fabs(x), ya = fabs(y);
    const double w = fmax(xa, ya), z = fmin(xa, ya);
    if (z == 0.0) return w;
    const double q = z / w;
    return w * sqrt(1.0 + q * q);
}
// ?larfg's range of plain squares: outside of it the norm is taken in scaled form (?nrm2 / ?lapy2) and |beta| < safmin
// triggers LAPACK's rescaling loop.  safmin = dlamch('S') / dlamch('E') = 2^-969.
constexpr double LARFG_SSQ_LO = 1e-280, LARFG_TOT_HI = 1e280;
constexpr double LARFG_SAFMIN = 2.2250738585072014e-308 / 1.1102230246251565e-16;

// (in the plain range only: no special-case branches)
__device__ __forceinline__ double strip_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);               // 1 - x y^2
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);                            // y (1 + e/2 + 3 e^2/8)
}
__device__ __forceinline__ double strip_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);                            // y (1 + e + e^2)
}

// One Householder step.  The register tile is kept ROTATED so that the pivot column always sits in
// register column SB-1 (static register indices, one loop body for every step: the fully unrolled
// variant was instruction-cache bound).  At step s (0-based) register column q holds
//   q <  s : the finished reflector of local column SB-s+q,
//   q >= s : the still-to-be-updated local column q-s      (q == SB-1: the pivot, local column SB-1-s).
template <int SB, int RPT>
__device__ __forceinline__ void strip_step(double (&a)[RPT][SB], const int (&rr)[RPT], StripSmem<SB> &sm, const int ps,
                                           const int rpc, const int tid, const uint32_t cta, const uint32_t nc, const int s) {
    constexpr int NW = STRIP_THREADS / 32;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int P = SB - 1;                        // register column of the pivot
    const int par = s & 1;
    const int warp = tid >> 5, lane = tid & 31;
    const int C = SB - 1 - s;                        // local column being eliminated
    const int mu = ps - 1 - s;                       // its pivot row
    // ---- 1. this thread's part of d_q = sum_{r < mu} x_r a_q[r], and the pivot row -----------------
    double x[RPT], d[SB];
#pragma unroll
    for (int j = 0; j < SB; ++j) d[j] = 0.0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        if (rr[i] == mu) {
#pragma unroll
            for (int j = 0; j < SB; ++j) sm.rowv_local[par][j] = a[i][j];
        }
        x[i] = (rr[i] < mu) ? a[i][P] : 0.0;
#pragma unroll
        for (int j = 0; j < SB; ++j) d[j] = fma(x[i], a[i][j], d[j]);
    }
    // ---- 2. warp reduction of SB values (recursive halving, fixed order) ----------------------------
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
    int vidx = 0;
    {
        int n = SB;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            if (n > 1) {
                n >>= 1;
                if (lane & o) vidx += n;
            }
        }
    }
    constexpr int LANES_PER_VAL = 32 / SB;           // lanes holding the same total (SB <= 32)
    if ((lane & (LANES_PER_VAL - 1)) == 0) sm.wpart[par][warp][vidx] = d[0];
    __syncthreads();
    // ---- 3. CTA partials -> every CTA of the cluster: st.async into their shared memory, completion
    //         counted on their transaction barrier (no cluster-wide hardware barrier on the critical path)
    const int owner = mu / rpc;
    if (tid < SB * (int)nc) {
        const int j = tid % SB;
        const uint32_t dst = tid / SB;
        double v = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < NW; ++w8) v += sm.wpart[par][w8][j];
        st_async_f64(&sm.red[par][cta][j], &sm.mbar[par], dst, v);
        if ((int)cta == owner) st_async_f64(&sm.rowv[par][j], &sm.mbar[par], dst, sm.rowv_local[par][j]);
    }
    mbar_wait_cluster(&sm.mbar[par], (uint32_t)((s >> 1) & 1));
    if (tid == 0) mbar_arrive_expect_tx(&sm.mbar[par], (uint32_t)((nc + 1) * SB * sizeof(double)));   // arm the next use
    // ---- 4. totals: every warp sums the cluster's partials itself (fixed pairwise order) ---------------
    double *tot = sm.totw[warp];
    if (lane < SB) {
        double v[STRIP_MAXC];
#pragma unroll
        for (int k = 0; k < STRIP_MAXC; ++k) v[k] = (k < (int)nc) ? sm.red[par][k][lane] : 0.0;
#pragma unroll
        for (int o = STRIP_MAXC / 2; o >= 1; o >>= 1)
#pragma unroll
            for (int k = 0; k < o; ++k) v[k] += v[k + o];
        tot[lane] = v[0];
    }
    __syncwarp();
    const double ssq = tot[P], alpha = sm.rowv[par][P];
    // ---- 5. ?larfg scalars ----------------------------------------------------------------------------
    double beta = alpha, tauc = 0.0, scale = 0.0;
    const bool plain = (ssq >= LARFG_SSQ_LO) && (fma(alpha, alpha, ssq) <= LARFG_TOT_HI);      // identical in every thread of the cluster
    if (plain) {
        // 1/sqrt and 1/x from the hardware seeds + one third-order FMA correction (below half an ulp before the final rounding,
        // the same arithmetic as the batched kernels): the IEEE sqrt and the two divisions were ~110 of the ~830 instructions
        // every warp issues per column, all of them on the serial chain
        const double totsq = fma(alpha, alpha, ssq);
        const double rinv = strip_rsqrt(totsq);
        const double nrm = totsq * rinv;
        beta = -copysign(nrm, alpha);
        tauc = fma(fabs(alpha), rinv, 1.0);            // (beta - alpha) / beta = 1 + |alpha| / nrm
        scale = strip_rcp(alpha - beta);
    } else {
        // LAPACK ?larfg, literally (what the reference reaches through LAPACK.geqlf!, src/ql.jl:72): the tail's norm in
        // scaled form, beta by ?lapy2, the safmin rescaling loop.  The dot products with the other columns are redone with
        // the tail scaled by its largest entry, so that neither the squares nor the products leave the exponent range.
        double am[1] = {0.0};
#pragma unroll
        for (int i = 0; i < RPT; ++i)
            if (rr[i] < mu) am[0] = fmax(am[0], fabs(a[i][P]));
        cluster_allreduce_slow<SB, 1, true>(am, sm, tid, nc);
        const double amax = sm.sl_tot[0];
        __syncthreads();                                   // sl_tot is rewritten by the next reduction
        if (amax > 0.0) {                                  // (zero tail: H = I, tau = 0, beta = alpha -- the defaults)
            double dv[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) dv[j] = 0.0;
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                x[i] = (rr[i] < mu) ? a[i][P] / amax : 0.0;
#pragma unroll
                for (int j = 0; j < SB; ++j) dv[j] = fma(x[i], (j == P) ? x[i] : a[i][j], dv[j]);
            }
            cluster_allreduce_slow<SB, SB, false>(dv, sm, tid, nc);
            const double ssq_s = sm.sl_tot[P];             // sum of (x_i / amax)^2
            double am_s = amax, al = alpha;
            double xnorm = am_s * sqrt(ssq_s);
            double be = -copysign(dlapy2_dev(al, xnorm), al);
            int knt = 0;
            if (fabs(be) < LARFG_SAFMIN) {
                const double rsafmn = 1.0 / LARFG_SAFMIN;
                do {
                    ++knt;
                    am_s *= rsafmn;
                    be *= rsafmn;
                    al *= rsafmn;
                } while (fabs(be) < LARFG_SAFMIN && knt < 20);
                xnorm = am_s * sqrt(ssq_s);
                be = -copysign(dlapy2_dev(al, xnorm), al);
            }
            tauc = (be - al) / be;
            const double f = am_s * (1.0 / (al - be));     // v_i = (x_i / amax) f,  |f| <= 1
            for (int q = 0; q < knt; ++q) be *= LARFG_SAFMIN;
            beta = be;
#pragma unroll
            for (int i = 0; i < RPT; ++i) x[i] *= f;       // x now holds the finished reflector tail
            scale = 1.0;
            if (lane < SB) tot[lane] = sm.sl_tot[lane] * f;      // v(tail) . a_j
            __syncwarp();
        }
        __syncthreads();
    }
    // ---- 6. rank-1 update of the unfinished columns (q >= s); the pivot becomes v (tail) and beta, and
    //         the tile rotates by one register column ----------------------------------------------------
    double vi[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) vi[i] = (rr[i] < mu) ? x[i] * scale : ((rr[i] == mu) ? 1.0 : 0.0);
    double newp[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) newp[i] = (rr[i] < mu) ? vi[i] : ((rr[i] == mu) ? beta : a[i][P]);
    // the update multipliers once per warp (lane j: column j) instead of once per thread and column, in place of the totals
    // (v_f . v_C of the finished columns is taken from them first)
    if (lane < SB) {
        const double w = fma(tot[lane], scale, sm.rowv[par][lane]);
        if (cta == 0 && warp == 0 && lane < s) sm.Gs[SB - s + lane][C] = w;
        if (cta == 0 && tid == 0) sm.taus[C] = tauc;
        tot[lane] = (lane >= s) ? -tauc * w : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int j = P - 1; j >= 0; --j) {
        const double coef = tot[j];
#pragma unroll
        for (int i = 0; i < RPT; ++i) a[i][j + 1] = fma(coef, vi[i], a[i][j]);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) a[i][0] = newp[i];
    __syncwarp();      // this warp's totals are rewritten in the next step
}

template <int SB, int RPT>
__global__ void __launch_bounds__(STRIP_THREADS, 1) strip_kernel(const StripParams p) {
    __shared__ StripSmem<SB> sm;
    double(&Gs)[SB][SB + 1] = sm.Gs;
    double(&Ts)[SB][SB + 1] = sm.Ts;
    double(&taus)[SB] = sm.taus;
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
    if (tid == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&sm.mbar[0], (uint32_t)((nc + 1) * SB * sizeof(double)));
        mbar_arrive_expect_tx(&sm.mbar[1], (uint32_t)((nc + 1) * SB * sizeof(double)));
    }
    cluster_sync_all();          // every CTA is resident and its barriers are armed before any remote store

#pragma unroll 1
    for (int s = 0; s < p.w; ++s) strip_step<SB, RPT>(a, rr, sm, p.ps, p.rpc, tid, cta, nc, s);
    // after w rotations register column q holds local column (q - w) mod SB

    // ---- write back the factors and the clean reflectors ----------------------------------------------
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rr[i];
        if (r != NOROW) {
#pragma unroll
            for (int q = 0; q < SB; ++q) {
                const int j = (q - p.w) & (SB - 1);          // local column held by register column q
                if (j >= off) {
                    const int mu = p.ps - (SB - j);
                    p.A[r + (int64_t)(j - off) * p.lda] = a[i][q];
                    p.Vw[r + (int64_t)(j - off) * p.ldv] = (r < mu) ? a[i][q] : ((r == mu) ? 1.0 : 0.0);
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

// W2 (rows x w, ldw) = -(sum_{s<S} part[s*stride + i + l*ldp]) * T (w x w, lower, ldt), w <= 16.
// blockDim = (16, 16, 4): x = local row (coalesced loads), y = column l, z = slice of the S partials
// (each slice sums a fixed quarter in a fixed order, the quarters are added pairwise).
__global__ void __launch_bounds__(1024) reduce_mult_t_kernel(const double *__restrict__ part, int S, int64_t stride, int rows, int w,
                                                            int64_t ldp, const double *__restrict__ T, int ldt,
                                                            double *__restrict__ W2, int64_t ldw) {
    __shared__ double slice[4][16][17];
    __shared__ double Tsm[16][17];
    const int ri = threadIdx.x, l = threadIdx.y, zs = threadIdx.z;
    const int i = blockIdx.x * 16 + ri;
    if (zs == 0) Tsm[ri][l] = (ri < w && l < w && ri >= l) ? T[ri + l * ldt] : 0.0;      // Tsm[row][col]
    double s = 0.0;
    if (i < rows && l < w) {
        const int per = (S + 3) / 4;
        const int z0 = zs * per, z1 = (z0 + per < S) ? z0 + per : S;
        const double *src = part + i + (int64_t)l * ldp;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int z = z0;
        for (; z + 4 <= z1; z += 4) {
            s0 += src[(int64_t)z * stride];
            s1 += src[(int64_t)(z + 1) * stride];
            s2 += src[(int64_t)(z + 2) * stride];
            s3 += src[(int64_t)(z + 3) * stride];
        }
        for (; z < z1; ++z) s0 += src[(int64_t)z * stride];
        s = (s0 + s1) + (s2 + s3);
    }
    slice[zs][ri][l] = s;
    __syncthreads();
    if (zs == 0) slice[0][ri][l] = (slice[0][ri][l] + slice[1][ri][l]) + (slice[2][ri][l] + slice[3][ri][l]);
    __syncthreads();
    if (zs == 0 && i < rows && l < w) {
        double o = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) o = fma(slice[0][ri][q], Tsm[q][l], o);
        W2[i + (int64_t)l * ldw] = -o;
    }
}

// T (ib x ib, lower triangular, zero above) from G = V'V (ib x ib, only the strict lower part is read)
// and tau; T(j,j) = tau_j, T(j+1:,j) = -tau_j * T(j+1:,j+1:) * G(j+1:,j), j = ib-1 ... 0 (?larft
// backward/columnwise).  One CTA, in place in shared memory, blocked by 16:
//   phase 1: the 16x16 diagonal blocks by the column recurrence, one warp per block, all in parallel;
//   phase 2: block columns right to left, T(below,b) = -(T(below,below) * G(below,b)) * T(b,b).
__global__ void __launch_bounds__(256, 1) larft_kernel(const double *__restrict__ G, int64_t ldg, const double *__restrict__ tau,
                                                      int ib, double *__restrict__ T, int64_t ldt) {
    extern __shared__ __align__(16) double sm[];
    double *W = sm;                                   // lower triangle of ib x ib, packed by columns (66 KB at ib = 128:
                                                      // the CTA fits beside one trailing-update GEMM CTA on an SM)
    auto PW = [&](int i, int j) -> double & { return W[j * ib - (j * (j - 1)) / 2 + (i - j)]; };      // i >= j
    double *ts = sm + (size_t)ib * (ib + 1) / 2;      // tau
    double *Ys = ts + 128;                            // (<=112) x 17 scratch
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < ib * ib; idx += blockDim.x) {
        const int i = idx % ib, j = idx / ib;
        if (i >= j) PW(i, j) = (i > j) ? G[i + (int64_t)j * ldg] : 0.0;
    }
    if (tid < ib) ts[tid] = tau[tid];
    __syncthreads();
    const int nblk = (ib + 15) / 16;
    // ---- phase 1: diagonal blocks ----
    for (int b = warp; b < nblk; b += 8) {
        const int r0 = b * 16, r1 = (r0 + 16 < ib) ? r0 + 16 : ib;
        const int i = r0 + lane;
        for (int j = r1 - 1; j >= r0; --j) {
            double s = 0.0;
            if (lane < 16 && i < r1 && i > j)
                for (int l = j + 1; l <= i; ++l) s = fma(PW(i, l), PW(l, j), s);
            __syncwarp();
            if (lane < 16 && i < r1) {
                if (i > j) PW(i, j) = -ts[j] * s;
                else if (i == j) PW(i, j) = ts[j];
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- phase 2: off-diagonal blocks, block column b from the right ----
    const int ri = tid >> 1, cg = (tid & 1) * 8;      // thread = (row below, group of 8 columns)
    for (int b = nblk - 2; b >= 0; --b) {
        const int c0 = b * 16, R0 = c0 + 16, R = ib - R0;
        double y[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) y[q] = 0.0;
        if (ri < R) {
            const int i = R0 + ri;
            for (int l = R0; l <= i; ++l) {          // T(i,l) * G(l, c0+cg+q)
                const double t = PW(i, l);
#pragma unroll
                for (int q = 0; q < 8; ++q) y[q] = fma(t, PW(l, c0 + cg + q), y[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) Ys[ri * 17 + cg + q] = y[q];
        }
        __syncthreads();
        if (ri < R) {
            const int i = R0 + ri;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int col = cg + q;               // Z(i,col) = -sum_{c >= col} Y(i,c) T_bb(c,col)
                double z = 0.0;
                for (int c = col; c < 16; ++c) z = fma(Ys[ri * 17 + c], PW(c0 + c, c0 + col), z);
                PW(i, c0 + col) = -z;
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < ib * ib; idx += blockDim.x) {
        const int r = idx % ib, c = idx / ib;
        T[r + (int64_t)c * ldt] = (r >= c) ? PW(r, c) : 0.0;
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

// ---------------------------------------------------------------------------------------------
// Strips taller than one cluster can hold (ps > 16 CTAs x 256 threads x 4 rows): the same column-by-column
// Householder step with the strip left in global memory -- two grid kernels per column (partial dots, then
// scalars + rank-1 update; every block re-sums the partials in the same fixed order).  Correctness path for very
// tall matrices (ql! of 65536 x 1024 and the like); the strip's T comes from its Gram matrix afterwards.
// ---------------------------------------------------------------------------------------------
constexpr int TALL_THREADS = 256;
constexpr int TALL_SB = 16;

// Partial sums of one block of rows, taken with the tail SCALED by the block's largest entry (amax_b, stored behind the
// partials): part[b][j] = sum_i (x_i / amax_b) a_ij for the columns j < C still to be updated, part[b][C] = sum_i
// (x_i / amax_b)^2.  Neither the squares nor the products can leave the exponent range (LAPACK ?nrm2 / ?larfg semantics
// for matrices scaled far away from 1).
__global__ void __launch_bounds__(TALL_THREADS) tall_dots_kernel(const double *__restrict__ A, int64_t lda, int mu, int C, int w,
                                                                 double *__restrict__ part, int rpb) {
    __shared__ double red[TALL_THREADS / 32][TALL_SB];
    __shared__ double mx[TALL_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * rpb, r1 = min(r0 + rpb, mu);
    double am = 0.0;
    for (int r = r0 + tid; r < r1; r += TALL_THREADS) am = fmax(am, fabs(A[r + (int64_t)C * lda]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
    if (lane == 0) mx[warp] = am;
    __syncthreads();
    am = mx[0];
    for (int q = 1; q < TALL_THREADS / 32; ++q) am = fmax(am, mx[q]);
    double d[TALL_SB];
#pragma unroll
    for (int j = 0; j < TALL_SB; ++j) d[j] = 0.0;
    if (am > 0.0) {
        for (int r = r0 + tid; r < r1; r += TALL_THREADS) {
            const double x = A[r + (int64_t)C * lda] / am;
#pragma unroll
            for (int j = 0; j < TALL_SB; ++j)
                if (j <= C) d[j] = fma(x, (j == C) ? x : A[r + (int64_t)j * lda], d[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < TALL_SB; ++j) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) d[j] += __shfl_xor_sync(0xffffffffu, d[j], o);
        if (lane == 0) red[warp][j] = d[j];
    }
    __syncthreads();
    if (tid < TALL_SB) {
        double t = 0.0;
        for (int q = 0; q < TALL_THREADS / 32; ++q) t += red[q][tid];
        part[(int64_t)blockIdx.x * TALL_SB + tid] = t;
        if (blockIdx.x == 0) part[(int64_t)gridDim.x * TALL_SB + tid] = (tid < w) ? A[mu + (int64_t)tid * lda] : 0.0;   // the pivot row
    }
    if (tid == 0) part[(int64_t)(gridDim.x + 1) * TALL_SB + blockIdx.x] = am;
}

__global__ void __launch_bounds__(TALL_THREADS) tall_update_kernel(double *__restrict__ A, int64_t lda, int mu, int C, int w,
                                                                   const double *__restrict__ part, double *__restrict__ tau_out,
                                                                   double *__restrict__ Vw, int64_t ldv, int p, int rpb) {
    __shared__ double tot[TALL_SB], rowv[TALL_SB], coef[TALL_SB], sc[2];
    const int tid = threadIdx.x;
    const int nblk = gridDim.x;
    const double *amaxs = part + (int64_t)(nblk + 1) * TALL_SB;
    if (tid < TALL_SB) {
        double amax = 0.0;
        for (int b = 0; b < nblk; ++b) amax = fmax(amax, amaxs[b]);
        double t = 0.0;
        if (amax > 0.0) {
            for (int b = 0; b < nblk; ++b) {
                const double wb = amaxs[b] / amax;                    // <= 1
                t += part[(int64_t)b * TALL_SB + tid] * (tid == C ? wb * wb : wb);
            }
        }
        tot[tid] = t;                                                 // x . a_j / amax  (j < C),  |x|^2 / amax^2  (j == C)
        rowv[tid] = part[(int64_t)nblk * TALL_SB + tid];
        if (tid == 0) sc[0] = amax;
    }
    __syncthreads();
    if (tid == 0) {
        // ?larfg (LAPACK, what src/ql.jl:72 runs): scaled norm, ?lapy2, safmin loop; v_i = (x_i / amax) f
        const double amax = sc[0], alpha = rowv[C];
        double beta = alpha, tauc = 0.0, f = 0.0;
        if (amax > 0.0) {
            double am_s = amax, al = alpha;
            double xnorm = am_s * sqrt(tot[C]);
            double be = -copysign(dlapy2_dev(al, xnorm), al);
            int knt = 0;
            if (fabs(be) < LARFG_SAFMIN) {
                const double rsafmn = 1.0 / LARFG_SAFMIN;
                do {
                    ++knt;
                    am_s *= rsafmn;
                    be *= rsafmn;
                    al *= rsafmn;
                } while (fabs(be) < LARFG_SAFMIN && knt < 20);
                xnorm = am_s * sqrt(tot[C]);
                be = -copysign(dlapy2_dev(al, xnorm), al);
            }
            tauc = (be - al) / be;
            f = am_s * (1.0 / (al - be));
            for (int q = 0; q < knt; ++q) be *= LARFG_SAFMIN;
            beta = be;
        }
        sc[1] = f;
        for (int j = 0; j < TALL_SB; ++j) coef[j] = (j < C) ? -tauc * fma(tot[j], f, rowv[j]) : 0.0;
        if (blockIdx.x == 0) {
            *tau_out = tauc;
            A[mu + (int64_t)C * lda] = beta;
            for (int j = 0; j < C; ++j) A[mu + (int64_t)j * lda] = rowv[j] + coef[j];      // a_j[mu] -= tau * s_j (v[mu] = 1)
            Vw[mu + (int64_t)C * ldv] = 1.0;
        }
    }
    __syncthreads();
    const double amax = sc[0], f = sc[1];
    const int r0 = blockIdx.x * rpb;
    for (int r = r0 + tid; r < r0 + rpb && r < p; r += TALL_THREADS) {
        if (r < mu) {
            const double v = (amax > 0.0) ? (A[r + (int64_t)C * lda] / amax) * f : 0.0;
            A[r + (int64_t)C * lda] = v;
            Vw[r + (int64_t)C * ldv] = v;
#pragma unroll
            for (int j = 0; j < TALL_SB; ++j)
                if (j < C) A[r + (int64_t)j * lda] = fma(coef[j], v, A[r + (int64_t)j * lda]);
        } else if (r > mu) {
            Vw[r + (int64_t)C * ldv] = 0.0;
        }
    }
}

static int strip_factor_tall(qlb200_ctx *h, cudaStream_t st, double *A, int64_t lda, int ps, int w, double *tau, double *Vw, int64_t ldv,
                             int p, double *T, int ldt) {
    const int rpb = 2048;
    const int nblk = (p + rpb - 1) / rpb;
    int rc = qlb_reserve(h, &h->tall, &h->tall_bytes, ((size_t)(nblk + 1) * TALL_SB + nblk + 256 + 64) * sizeof(double));
    if (rc) return rc;
    double *part = (double *)h->tall, *G = part + (size_t)(nblk + 1) * TALL_SB + ((nblk + 1) & ~1);
    for (int s = 0; s < w; ++s) {
        const int C = w - 1 - s, mu = ps - 1 - s;
        tall_dots_kernel<<<nblk, TALL_THREADS, 0, st>>>(A, lda, mu, C, w, part, rpb);
        QLB_LAUNCH_CHECK(h);
        tall_update_kernel<<<nblk, TALL_THREADS, 0, st>>>(A, lda, mu, C, w, part, tau + C, Vw, ldv, p, rpb);
        QLB_LAUNCH_CHECK(h);
    }
    // T of the strip from the Gram matrix of its clean reflectors (?larft backward/columnwise)
    rc = qlb_gemm(h, st, 2, true, true, w, w, p, 1.0, Vw, ldv, Vw, ldv, 0.0, G, w, 1, 0);
    if (rc) return rc;
    return qlb_larft(h, st, G, w, tau, w, T, ldt);
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

// Largest strip (rows) the cluster kernel can hold: 16 CTAs x STRIP_THREADS x 4 rows.
int qlb_strip_max_rows() { return qlb::STRIP_MAXC * qlb::STRIP_THREADS * 4; }

int qlb_strip_factor(qlb200_ctx *h, cudaStream_t st, double *A, int64_t lda, int ps, int w, double *tau, double *Vw,
                     int64_t ldv, int p, double *T, int ldt) {
    using namespace qlb;
    if (w < 1 || w > 16) return qlb_fail(h, -2, "strip width must be in [1,16]%s");
    if (ps > qlb_strip_max_rows()) return qlb::strip_factor_tall(h, st, A, lda, ps, w, tau, Vw, ldv, p, T, ldt);
    // rows per thread (smallest of 1/2/4 that fits 16 CTAs, or the "strip_rpt" option), then CTAs per
    // cluster (power of two <= 16)
    int rpt = (ps + STRIP_THREADS * STRIP_MAXC - 1) / (STRIP_THREADS * STRIP_MAXC);
    rpt = rpt <= 1 ? 1 : (rpt <= 2 ? 2 : 4);
    if (h->opt_strip_rpt > rpt) rpt = (int)h->opt_strip_rpt;
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
    dim3 block(16, 16, 4);
    qlb::reduce_mult_t_kernel<<<(rows + 15) / 16, block, 0, st>>>(part, S, stride, rows, w, ldp, T, ldt, W2, ldw);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

int qlb_larft(qlb200_ctx *h, cudaStream_t st, const double *G, int64_t ldg, const double *tau, int ib, double *T,
              int64_t ldt) {
    if (ib <= 0) return 0;
    if (ib > 128) return qlb_fail(h, -2, "larft: block wider than 128%s");
    const size_t smem = ((size_t)ib * (ib + 1) / 2 + 128 + 112 * 17) * sizeof(double);
    if (!h->larft_attr_done) {
        QLB_CUDA(h, cudaFuncSetAttribute(qlb::larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (128 * 129 / 2 + 128 + 112 * 17) * 8));
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
