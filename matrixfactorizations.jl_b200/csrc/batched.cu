// batched.cu -- K6: batched small square Householder QL, Q-apply and square solve (sm_100a).
//
// Per-matrix semantics = the reference's ql! for BlasFloat (LAPACK ?geqlf -> ?geql2 + ?larfg sign
// conventions, reference src/ql.jl:72 with the unblocked spec at src/ql.jl:56-68), lmul!(Q,B) /
// lmul!(Q',B) (src/ql.jl:321-347 / 350-377) and the square ldiv!(F::QL,B) (src/ql.jl:440-447,467).
//
// Factorization kernel: a group of LPM = N/CPL lanes owns one matrix; every lane keeps CPL whole
// columns in registers, so the reflector dot products v^T a_j and the rank-1 update are lane-local
// FMAs (no shuffles); only the pivot column is broadcast (through shared memory) and two scalars
// are shuffled per step.  Matrices are staged global -> shared by TMA 1-D bulk copies (one per
// column, padded pitch => conflict-free column reads), prefetched one warp-batch ahead; factors go
// back with 16-byte stores straight from registers.  HBM traffic = read A + write factors + write
// tau = (2 n^2 + n) * sizeof(T) bytes per matrix, the compulsory minimum.
#include "common.cuh"
#include "internal.h"

namespace qlb {

template <typename T> struct Vec16;
template <> struct Vec16<double> { using type = double2; static constexpr int E = 2; };
template <> struct Vec16<float>  { using type = float4;  static constexpr int E = 4; };

template <typename T> struct Lim;
template <> struct Lim<double> { static constexpr double LO = 1e-280, HI = 1e280; };
template <> struct Lim<float>  { static constexpr float  LO = 1e-24f, HI = 1e30f; };

__device__ __forceinline__ double fast_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float  fast_rsqrt(float x)  { return rsqrtf(x); }
__device__ __forceinline__ double fast_rcp(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float  fast_rcp(float x)  { return __frcp_rn(x); }
__device__ __forceinline__ double t_abs(double x) { return fabs(x); }
__device__ __forceinline__ float  t_abs(float x)  { return fabsf(x); }
__device__ __forceinline__ double t_fma(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ float  t_fma(float a, float b, float c)  { return fmaf(a, b, c); }
__device__ __forceinline__ double t_copysign(double a, double b) { return copysign(a, b); }
__device__ __forceinline__ float  t_copysign(float a, float b)  { return copysignf(a, b); }
__device__ __forceinline__ double t_scalbn(double a, int e) { return scalbn(a, e); }
__device__ __forceinline__ float  t_scalbn(float a, int e)  { return scalbnf(a, e); }
__device__ __forceinline__ int t_ilogb(double a) { return ilogb(a); }
__device__ __forceinline__ int t_ilogb(float a)  { return ilogbf(a); }

// Volatile 16-byte shared loads: keep the pivot-column broadcast reads where they are written (two
// passes re-read x from shared memory instead of pinning N more registers per lane).
__device__ __forceinline__ double2 lds16(const double *p) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ float4 lds16(const float *p) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_u32(p)));
    return r;
}

template <typename T, int N, int CPL>
struct BatchedCfg {
    static constexpr int E = Vec16<T>::E;            // elements per 16-byte vector
    static constexpr int LPM = N / CPL;              // lanes per matrix
    static constexpr int G = 32 / LPM;               // matrices per warp
    static constexpr int PITCH = N + E;              // padded column pitch (elements) -> no bank conflicts
    static constexpr int NCOLS = G * N;              // columns staged per warp-batch
    static constexpr int XPITCH = N + E;             // per-group pivot-column broadcast buffer
    static constexpr int BUF_BYTES = NCOLS * PITCH * (int)sizeof(T);
    static constexpr int X_BYTES = G * XPITCH * (int)sizeof(T);
    static constexpr int WARP_BYTES = BUF_BYTES + X_BYTES + 16;   // + mbarrier slot
};

// Rare, warp-uniform check shared by all pivots (kept out of line so that it costs no code in the hot
// loop).  ?larfg has two special exits.  A tail that is EXACTLY zero gives tau = 0 / H = I: handled
// in line by the caller (returns true).  A tail whose squares under/overflow needs LAPACK's scaled
// norm and safmin rescaling loop: such matrices are flagged, left untouched by the fast kernel and
// re-factored by geql2_generic_kernel, which follows ?larfg literally.
template <typename T>
__device__ __noinline__ bool ql_rare_zero_tail(const T *xg, int R) {
    T amax = T(0);
    for (int i = 0; i < R; ++i) amax = fmax(amax, t_abs(xg[i]));
    return amax == T(0);
}

// 1/sqrt(x) and 1/x for x in the normal range: hardware approximation (MUFU.RSQ64H / RCP64H) plus
// Newton steps in FMA arithmetic -- no special-case branches (callers guarantee the range).
__device__ __forceinline__ double nr_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double e = fma(-hx * y, y, 0.5);      // 0.5 - 0.5 x y^2
        y = fma(y, e, y);
    }
    return y;
}
__device__ __forceinline__ float nr_rsqrt(float x) { return rsqrtf(x) ; }
__device__ __forceinline__ double nr_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double e = fma(-x, y, 1.0);
        y = fma(y, e, y);
    }
    return y;
}
__device__ __forceinline__ float nr_rcp(float x) { return __frcp_rn(x); }

// One Householder QL step at compile-time pivot row/column R (0-based), LAPACK ?larfg conventions.
template <typename T, int N, int CPL, int R>
__device__ __forceinline__ void ql_step(T (&a)[CPL][N], T (&mytau)[CPL], T (&myscale)[CPL], T *xg, int jl, bool &bad) {
    using V = typename Vec16<T>::type;
    constexpr int E = Vec16<T>::E;
    constexpr int LPM = N / CPL;
    constexpr int OWN_LANE = R % LPM, OWN_SLOT = R / LPM;
    constexpr unsigned FULL = 0xffffffffu;
    if constexpr (R == 0) {
        // length-1 reflector: tau = 0, beta = alpha (H = I)
        if (jl == 0) { mytau[0] = T(0); myscale[0] = T(0); }
        return;
    } else {
        constexpr int NV = (R + E - 1) / E;      // 16-byte vectors covering the tail x[0..R-1]
        // 1. owner publishes the pivot column's tail
        if (jl == OWN_LANE) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                V q;
                T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
                for (int e = 0; e < E; ++e) qp[e] = a[OWN_SLOT][v * E + e];
                reinterpret_cast<V *>(xg)[v] = q;
            }
        }
        __syncwarp();
        // 2. tail dot products d_c = sum_{i<R} x_i a[c][i] for every slot that still has live columns
        //    (four fixed accumulation chains: the summation order never depends on the data)
        T d[CPL], d1[CPL], d2[CPL], d3[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) { d[c] = T(0); d1[c] = T(0); d2[c] = T(0); d3[c] = T(0); }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            V q = lds16(xg + v * E);
            const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = v * E + e;
                if (i < R) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c)
                        if (c * LPM <= R) {
                            if ((i & 3) == 0) d[c] = t_fma(qp[e], a[c][i], d[c]);
                            else if ((i & 3) == 1) d1[c] = t_fma(qp[e], a[c][i], d1[c]);
                            else if ((i & 3) == 2) d2[c] = t_fma(qp[e], a[c][i], d2[c]);
                            else d3[c] = t_fma(qp[e], a[c][i], d3[c]);
                        }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CPL; ++c) d[c] = (d[c] + d1[c]) + (d2[c] + d3[c]);
        T ssq = __shfl_sync(FULL, d[OWN_SLOT], OWN_LANE, LPM);
        T alpha = __shfl_sync(FULL, a[OWN_SLOT][R], OWN_LANE, LPM);
        T tot = t_fma(alpha, alpha, ssq);
        bool zero_tail = false;
        const bool safe = (ssq >= Lim<T>::LO) && (tot <= Lim<T>::HI);
        if (__any_sync(FULL, !safe)) {
            zero_tail = ql_rare_zero_tail<T>(xg, R);
            if (!safe) {
                if (!zero_tail) bad = true;      // needs ?larfg's scaled path: redo in the generic kernel
                tot = T(1);
            } else {
                zero_tail = false;
            }
        }
        // 3. scalars: beta = -sign(alpha) ||[alpha;x]||, tau = (beta-alpha)/beta = 1 + |alpha|/||.||,
        //    scale = 1/(alpha-beta) = -1/(tau*beta)
        const T rinv = nr_rsqrt(tot);                    // 1/||.||
        T nrm = tot * rinv;
        nrm = t_fma(t_fma(-nrm, nrm, tot), T(0.5) * rinv, nrm);   // one Newton correction of the norm
        const T beta = -t_copysign(nrm, alpha);
        T tauk = t_fma(t_abs(alpha), rinv, T(1));
        T scale = -nr_rcp(tauk * beta);
        if (zero_tail) { tauk = T(0); scale = T(0); }
        // 4. owner bookkeeping (scaling of v deferred to the end of the factorization)
        if (jl == OWN_LANE) {
            mytau[OWN_SLOT] = tauk;
            myscale[OWN_SLOT] = scale;                 // v_i = x_i / (alpha - beta)
            if (!zero_tail) a[OWN_SLOT][R] = beta;
        }
        // 5. rank-1 update of the live columns left of the pivot: a_j -= tau (v^T a_j) v
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            if (c * LPM < R) {
                const bool live = (jl + c * LPM) < R;
                const T s = t_fma(d[c], scale, a[c][R]);
                const T t = live ? tauk * s : T(0);
                a[c][R] -= t;
                d[c] = -t * scale;                      // multiplier on raw x
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            V q = lds16(xg + v * E);
            const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = v * E + e;
                if (i < R) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c)
                        if (c * LPM < R) a[c][i] = t_fma(d[c], qp[e], a[c][i]);
                }
            }
        }
        __syncwarp();     // xg is rewritten by the next step's owner
    }
}

template <typename T, int N, int CPL, int R>
struct StepLoop {
    static __device__ __forceinline__ void run(T (&a)[CPL][N], T (&mytau)[CPL], T (&myscale)[CPL], T *xg, int jl, bool &bad) {
        ql_step<T, N, CPL, R>(a, mytau, myscale, xg, jl, bad);
        if constexpr (R > 0) StepLoop<T, N, CPL, R - 1>::run(a, mytau, myscale, xg, jl, bad);
    }
};

template <typename T, int N, int CPL, int WARPS, int MINB, bool USE_TMA>
__global__ void __launch_bounds__(WARPS * 32, MINB)
geqlf_batched_kernel(T *__restrict__ A, int64_t lda, int64_t strideA, T *__restrict__ tau, int64_t strideTau,
                     int64_t batch, int *__restrict__ redo_count, int *__restrict__ redo_list) {
    using Cfg = BatchedCfg<T, N, CPL>;
    using V = typename Vec16<T>::type;
    constexpr int E = Cfg::E, LPM = Cfg::LPM, G = Cfg::G, PITCH = Cfg::PITCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LPM, jl = lane % LPM;
    unsigned char *wbase = smem_raw + (size_t)warp * Cfg::WARP_BYTES;
    T *buf = reinterpret_cast<T *>(wbase);
    T *xg = reinterpret_cast<T *>(wbase + Cfg::BUF_BYTES) + g * Cfg::XPITCH;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + Cfg::BUF_BYTES + Cfg::X_BYTES);

    const int64_t nwb = (batch + G - 1) / G;                 // warp-batches of G matrices
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;

    if constexpr (USE_TMA) {
        if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
        __syncwarp();
    }
    auto issue = [&](int64_t wb) {
        const int64_t m0 = wb * G;
        const int64_t nvalid = (batch - m0) < G ? (batch - m0) : G;
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(nvalid * N * N * sizeof(T)));
        const int64_t mi = m0 + g;
        if (mi < batch) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = jl + c * LPM;
                tma_load_1d(buf + (g * N + col) * PITCH, A + mi * strideA + (int64_t)col * lda, N * sizeof(T), bar);
            }
        }
    };
    uint32_t phase = 0;
    if constexpr (USE_TMA) {
        if (gw < nwb) issue(gw);
    }
    for (int64_t wb = gw; wb < nwb; wb += nw) {
        const int64_t mi = wb * G + g;
        const bool valid = mi < batch;
        T a[CPL][N];
        if constexpr (USE_TMA) {
            mbar_wait(bar, phase);
            phase ^= 1;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const V *src = reinterpret_cast<const V *>(buf + (g * N + jl + c * LPM) * PITCH);
#pragma unroll
                for (int v = 0; v < N / E; ++v) {
                    V q = src[v];
                    const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
                    for (int e = 0; e < E; ++e) a[c][v * E + e] = qp[e];
                }
            }
            __syncwarp();
            if (wb + nw < nwb) issue(wb + nw);            // prefetch the next warp-batch into the same buffer
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const T *src = A + (valid ? mi : 0) * strideA + (int64_t)(jl + c * LPM) * lda;
#pragma unroll
                for (int i = 0; i < N; ++i) a[c][i] = src[i];
            }
        }
        if (!valid) {
#pragma unroll
            for (int c = 0; c < CPL; ++c)
#pragma unroll
                for (int i = 0; i < N; ++i) a[c][i] = (i == jl + c * LPM) ? T(1) : T(0);
        }
        T mytau[CPL], myscale[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) { mytau[c] = T(0); myscale[c] = T(0); }

        bool bad = false;
        StepLoop<T, N, CPL, N - 1>::run(a, mytau, myscale, xg, jl, bad);

        // deferred scaling of the reflector tails: v_i = x_i / (alpha - beta), rows above the pivot only
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = jl + c * LPM;
#pragma unroll
            for (int i = 0; i < N - 1; ++i) a[c][i] *= (i < col) ? myscale[c] : T(1);
        }
        if (valid && bad && jl == 0) redo_list[atomicAdd(redo_count, 1)] = (int)mi;      // untouched; redone generically
        if (valid && !bad) {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = jl + c * LPM;
                T *dst = A + mi * strideA + (int64_t)col * lda;
                if constexpr (USE_TMA) {
#pragma unroll
                    for (int v = 0; v < N / E; ++v) {
                        V q;
                        T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
                        for (int e = 0; e < E; ++e) qp[e] = a[c][v * E + e];
                        reinterpret_cast<V *>(dst)[v] = q;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < N; ++i) dst[i] = a[c][i];
                }
                tau[mi * strideTau + col] = mytau[c];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Generic robust QL: one warp per matrix, matrix resident in shared memory, any 1 <= n <= 64.
// Follows LAPACK ?geql2 / ?larfg literally (scaled 2-norm, ?lapy2, the safmin rescaling loop), so it
// is the path for (a) matrices the register kernels flagged (squares under/overflow), and (b) sizes
// without a specialised kernel.  Matrices come from `list` (count = *count_ptr) or, when list is
// null, are simply 0 .. count_val-1.
// ---------------------------------------------------------------------------------------------
template <typename T> struct LarfgConst;
template <> struct LarfgConst<double> { static __device__ double safmin() { return 2.2250738585072014e-308 / 1.1102230246251565e-16; } };
template <> struct LarfgConst<float>  { static __device__ float  safmin() { return 1.17549435e-38f / 5.96046448e-08f; } };

template <typename T>
__device__ __forceinline__ T warp_nrm2(const T *x, int len, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    T amax = T(0);
    for (int i = lane; i < len; i += 32) amax = fmax(amax, t_abs(x[i]));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) amax = fmax(amax, __shfl_xor_sync(FULL, amax, o));
    if (amax == T(0)) return T(0);
    T s = T(0);
    for (int i = lane; i < len; i += 32) {
        const T q = x[i] / amax;
        s = t_fma(q, q, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    return amax * sqrt(s);
}
template <typename T>
__device__ __forceinline__ T t_lapy2(T x, T y) {
    const T xa = t_abs(x), ya = t_abs(y);
    const T w = fmax(xa, ya), z = fmin(xa, ya);
    if (z == T(0)) return w;
    const T q = z / w;
    return w * sqrt(T(1) + q * q);
}

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) geql2_generic_kernel(T *__restrict__ A, int64_t lda, int64_t strideA, T *__restrict__ tau,
                                                                  int64_t strideTau, int n, const int *__restrict__ count_ptr,
                                                                  int64_t count_val, const int *__restrict__ list) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int P = n + 1;                                     // column pitch
    T *S = reinterpret_cast<T *>(smem_raw) + (size_t)warp * (P * n + n);
    T *taus = S + P * n;
    const int64_t count = list ? (int64_t)*count_ptr : count_val;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
    const T safmin = LarfgConst<T>::safmin(), rsafmn = T(1) / safmin;
    for (int64_t it = gw; it < count; it += nw) {
        const int64_t mi = list ? (int64_t)list[it] : it;
        T *Am = A + mi * strideA;
        for (int idx = lane; idx < n * n; idx += 32) S[(idx / n) * P + (idx % n)] = Am[(int64_t)(idx / n) * lda + (idx % n)];
        __syncwarp();
        for (int k = n - 1; k >= 0; --k) {
            T *x = S + k * P;                                // column k: tail rows 0..k-1, pivot row k
            T tk = T(0);
            T xnorm = (k > 0) ? warp_nrm2<T>(x, k, lane) : T(0);
            if (xnorm != T(0)) {
                T alpha = x[k];
                T beta = -t_copysign(t_lapy2(alpha, xnorm), alpha);
                int knt = 0;
                if (t_abs(beta) < safmin) {
                    do {
                        ++knt;
                        __syncwarp();
                        for (int i = lane; i < k; i += 32) x[i] *= rsafmn;
                        beta *= rsafmn;
                        alpha *= rsafmn;
                    } while (t_abs(beta) < safmin && knt < 20);
                    __syncwarp();
                    xnorm = warp_nrm2<T>(x, k, lane);
                    beta = -t_copysign(t_lapy2(alpha, xnorm), alpha);
                }
                tk = (beta - alpha) / beta;
                const T sc = T(1) / (alpha - beta);
                __syncwarp();
                for (int i = lane; i < k; i += 32) x[i] *= sc;
                for (int j = 0; j < knt; ++j) beta *= safmin;
                __syncwarp();
                if (lane == 0) x[k] = beta;
                // apply H = I - tau v v' (v = [x(0:k-1); 1]) to columns j < k, one column per lane
                for (int j = lane; j < k; j += 32) {
                    T *c = S + j * P;
                    T w = c[k];
                    for (int i = 0; i < k; ++i) w = t_fma(x[i], c[i], w);
                    w *= tk;
                    c[k] -= w;
                    for (int i = 0; i < k; ++i) c[i] = t_fma(-w, x[i], c[i]);
                }
            }
            if (lane == 0) taus[k] = tk;
            __syncwarp();
        }
        for (int idx = lane; idx < n * n; idx += 32) Am[(int64_t)(idx / n) * lda + (idx % n)] = S[(idx / n) * P + (idx % n)];
        for (int i = lane; i < n; i += 32) tau[mi * strideTau + i] = taus[i];
        __syncwarp();
    }
}

template <typename T>
static int launch_geql2_generic(qlb200_ctx *h, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau, int n,
                                const int *count_ptr, int64_t count_val, const int *list) {
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * ((size_t)(n + 1) * n + n) * sizeof(T);
    auto kern = geql2_generic_kernel<T, WARPS>;
    QLB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = list ? (int64_t)h->sm_count : (count_val + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)h->sm_count * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(A, lda, strideA, tau, strideTau, n, count_ptr, count_val, list);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Q-apply and square solve: one warp per matrix.  Lane (jb, rg) = (lane % NR, lane / NR) owns
// RPL = N*NR/32 consecutive rows of column jb of B; reflector k's vector is read from the factors
// staged in shared memory, the dot product is finished with log2(32/NR) xor-shuffles.
// MODE 0: B <- Q B (k ascending, src/ql.jl:321-347); MODE 1: B <- Q' B (k descending,
// src/ql.jl:350-377); MODE 2: B <- L^{-1} Q' B (src/ql.jl:440-447,467).
// ---------------------------------------------------------------------------------------------
template <typename T, int N, int NR>
struct ApplyCfg {
    static constexpr int E = Vec16<T>::E;
    static constexpr int RG = 32 / NR;               // row groups
    static constexpr int RPL = N / RG;               // rows per lane
    static constexpr int PITCH = N + E;
    static constexpr int BUF_BYTES = N * PITCH * (int)sizeof(T);
    static constexpr int WARP_BYTES = BUF_BYTES + N * (int)sizeof(T) + 16;
};

template <typename T, int N, int NR, int MODE, int WARPS, bool USE_TMA>
__global__ void __launch_bounds__(WARPS * 32)
apply_batched_kernel(const T *__restrict__ A, int64_t lda, int64_t strideA, const T *__restrict__ tau,
                     int64_t strideTau, T *__restrict__ B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch,
                     int32_t *__restrict__ info_out) {
    using Cfg = ApplyCfg<T, N, NR>;
    constexpr int RPL = Cfg::RPL, PITCH = Cfg::PITCH;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb = lane % NR, rg = lane / NR, row0 = rg * RPL;
    unsigned char *wbase = smem_raw + (size_t)warp * Cfg::WARP_BYTES;
    T *buf = reinterpret_cast<T *>(wbase);
    T *taus = reinterpret_cast<T *>(wbase + Cfg::BUF_BYTES);
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + Cfg::BUF_BYTES + N * sizeof(T));
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
    if constexpr (USE_TMA) {
        if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
        __syncwarp();
    }
    uint32_t phase = 0;
    for (int64_t mi = gw; mi < batch; mi += nw) {
        const T *Am = A + mi * strideA;
        // stage the factors (and tau) of this matrix
        if constexpr (USE_TMA) {
            if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)(N * N * sizeof(T)));
            for (int col = lane; col < N; col += 32) tma_load_1d(buf + col * PITCH, Am + (int64_t)col * lda, N * sizeof(T), bar);
        } else {
            for (int idx = lane; idx < N * N; idx += 32) buf[(idx / N) * PITCH + (idx % N)] = Am[(int64_t)(idx / N) * lda + (idx % N)];
        }
        for (int i = lane; i < N; i += 32) taus[i] = tau[mi * strideTau + i];
        // this lane's slice of B: nrhs columns are processed NR at a time
        for (int c0 = 0; c0 < nrhs; c0 += NR) {
            const int col = c0 + jb;
            const bool cvalid = col < nrhs;
            T b[RPL];
            T *Bc = B + mi * strideB + (int64_t)(cvalid ? col : 0) * ldb + row0;
#pragma unroll
            for (int i = 0; i < RPL; ++i) b[i] = cvalid ? Bc[i] : T(0);
            if (c0 == 0) {
                if constexpr (USE_TMA) { mbar_wait(bar, phase); phase ^= 1; }
                __syncwarp();
            }
            // reflectors: column k (0-based) has pivot row k, tail rows 0..k-1
            for (int kk = 0; kk < N; ++kk) {
                const int k = (MODE == 0) ? kk : N - 1 - kk;
                const T *vk = buf + k * PITCH + row0;
                const T tk = taus[k];
                T v[RPL];
                T part = T(0);
#pragma unroll
                for (int i = 0; i < RPL; ++i) {
                    const int row = row0 + i;
                    v[i] = (row < k) ? vk[i] : (row == k ? T(1) : T(0));
                    part = t_fma(v[i], b[i], part);
                }
#pragma unroll
                for (int off = NR; off < 32; off <<= 1) part += __shfl_xor_sync(FULL, part, off);
                const T s = tk * part;
#pragma unroll
                for (int i = 0; i < RPL; ++i) b[i] = t_fma(-s, v[i], b[i]);
            }
            if constexpr (MODE == 2) {
                // forward substitution with L = tril(factors): x_c = b_c / L_cc; b_i -= L_ic x_c (i > c)
                int bad = 0;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const int orow = c % RPL, org = c / RPL;       // owner row-group / local row of row c
                    const T *lc = buf + c * PITCH + row0;
                    const T lcc = buf[c * PITCH + c];
                    if (lcc == T(0) && bad == 0) bad = c + 1;
                    T xc = T(0);
#pragma unroll
                    for (int i = 0; i < RPL; ++i)
                        if (i == orow) xc = b[i];
                    xc = __shfl_sync(FULL, xc, jb + NR * org) / lcc;
#pragma unroll
                    for (int i = 0; i < RPL; ++i) {
                        const int row = row0 + i;
                        if (row == c) b[i] = xc;
                        else if (row > c) b[i] = t_fma(-lc[i], xc, b[i]);
                    }
                }
                if (info_out != nullptr && lane == 0 && c0 == 0) info_out[mi] = bad;
            }
            if (cvalid) {
#pragma unroll
                for (int i = 0; i < RPL; ++i) Bc[i] = b[i];
            }
        }
        __syncwarp();    // all lanes are done with buf/taus before the next matrix is staged
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
template <typename T>
static bool tma_ok(const T *A, int64_t lda, int64_t strideA) {
    return ((uintptr_t)A % 16 == 0) && ((lda * sizeof(T)) % 16 == 0) && ((strideA * sizeof(T)) % 16 == 0);
}

template <typename T, int N, int CPL, int WARPS, int MINB>
static int launch_geqlf_batched(qlb200_ctx *h, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau,
                                int64_t batch) {
    using Cfg = BatchedCfg<T, N, CPL>;
    const size_t smem = (size_t)WARPS * Cfg::WARP_BYTES;
    const bool tma = tma_ok(A, lda, strideA);
    auto kern = tma ? geqlf_batched_kernel<T, N, CPL, WARPS, MINB, true> : geqlf_batched_kernel<T, N, CPL, WARPS, MINB, false>;
    QLB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    QLB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
    if (occ < 1) return qlb_fail(h, 1000 + (int)cudaErrorLaunchOutOfResources, "batched QL kernel does not fit on an SM%s");
    const int64_t nwb = (batch + Cfg::G - 1) / Cfg::G;
    int64_t blocks = (nwb + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)h->sm_count * occ;        // persistent: one resident wave
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    // flagged matrices (under/overflowing squares) are collected here and redone by the generic kernel
    if (batch >= ((int64_t)1 << 31)) return qlb_fail(h, -8, "batch too large%s");
    int rc = qlb_reserve(h, &h->redo, &h->redo_bytes, ((size_t)batch + 4) * sizeof(int));
    if (rc) return rc;
    int *redo_count = (int *)h->redo, *redo_list = redo_count + 4;
    QLB_CUDA(h, cudaMemsetAsync(redo_count, 0, sizeof(int), h->stream));
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(A, lda, strideA, tau, strideTau, batch, redo_count, redo_list);
    QLB_LAUNCH_CHECK(h);
    return launch_geql2_generic<T>(h, A, lda, strideA, tau, strideTau, N, redo_count, 0, redo_list);
}

template <typename T, int N, int NR, int MODE>
static int launch_apply_batched(qlb200_ctx *h, const T *A, int64_t lda, int64_t strideA, const T *tau,
                                int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch,
                                int32_t *info_out) {
    constexpr int WARPS = 4;
    using Cfg = ApplyCfg<T, N, NR>;
    const size_t smem = (size_t)WARPS * Cfg::WARP_BYTES;
    const bool tma = tma_ok(A, lda, strideA);
    auto kern = tma ? apply_batched_kernel<T, N, NR, MODE, WARPS, true> : apply_batched_kernel<T, N, NR, MODE, WARPS, false>;
    QLB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    QLB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
    if (occ < 1) return qlb_fail(h, 1000 + (int)cudaErrorLaunchOutOfResources, "batched apply kernel does not fit on an SM%s");
    int64_t blocks = (batch + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)h->sm_count * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch,
                                                           info_out);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T, int N, int MODE>
static int dispatch_apply_nr(qlb200_ctx *h, const T *A, int64_t lda, int64_t strideA, const T *tau, int64_t strideTau,
                             T *B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch, int32_t *info_out) {
    // NR = columns of B handled per pass: a power of two <= 8 with N*NR >= 32 rows*cols per warp
    constexpr int NRMIN = (32 / N) > 1 ? (32 / N) : 1;
    if (nrhs >= 8 || NRMIN >= 8) return launch_apply_batched<T, N, 8, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
    if (nrhs > 2 || NRMIN >= 4) return launch_apply_batched<T, N, 4, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
    if (nrhs > 1 || NRMIN >= 2) return launch_apply_batched<T, N, (NRMIN > 2 ? NRMIN : 2), MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
    return launch_apply_batched<T, N, (NRMIN > 1 ? NRMIN : 1), MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
}

template <typename T, int MODE>
static int dispatch_apply_n(qlb200_ctx *h, int64_t n, const T *A, int64_t lda, int64_t strideA, const T *tau,
                            int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch,
                            int32_t *info_out) {
    switch (n) {
        case 8: return dispatch_apply_nr<T, 8, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
        case 16: return dispatch_apply_nr<T, 16, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
        case 32: return dispatch_apply_nr<T, 32, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
    }
    return qlb_fail(h, -2, "batched apply: n must be 8, 16 or 32%s");
}

}  // namespace qlb

// ---------------------------------------------------------------------------------------------
// typed internal entry points used by capi.cu
// ---------------------------------------------------------------------------------------------
int qlb_geqlf_batched_f64(qlb200_ctx *h, int64_t n, double *A, int64_t lda, int64_t strideA, double *tau,
                          int64_t strideTau, int64_t batch) {
    switch (n) {
        case 8: return qlb::launch_geqlf_batched<double, 8, 1, 4, 1>(h, A, lda, strideA, tau, strideTau, batch);
        case 16: return qlb::launch_geqlf_batched<double, 16, 2, 4, 1>(h, A, lda, strideA, tau, strideTau, batch);
        case 32: return qlb::launch_geqlf_batched<double, 32, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch);
    }
    if (n >= 1 && n <= 64) return qlb::launch_geql2_generic<double>(h, A, lda, strideA, tau, strideTau, (int)n, nullptr, batch, nullptr);
    return qlb_fail(h, -2, "batched QL: n must be in [1,64]%s");
}
int qlb_geqlf_batched_f32(qlb200_ctx *h, int64_t n, float *A, int64_t lda, int64_t strideA, float *tau,
                          int64_t strideTau, int64_t batch) {
    switch (n) {
        case 8: return qlb::launch_geqlf_batched<float, 8, 1, 4, 1>(h, A, lda, strideA, tau, strideTau, batch);
        case 16: return qlb::launch_geqlf_batched<float, 16, 2, 4, 1>(h, A, lda, strideA, tau, strideTau, batch);
        case 32: return qlb::launch_geqlf_batched<float, 32, 2, 4, 1>(h, A, lda, strideA, tau, strideTau, batch);
    }
    if (n >= 1 && n <= 64) return qlb::launch_geql2_generic<float>(h, A, lda, strideA, tau, strideTau, (int)n, nullptr, batch, nullptr);
    return qlb_fail(h, -2, "batched QL: n must be in [1,64]%s");
}
// mode 0: Q*B, 1: Q'*B, 2: L^{-1} Q' B
int qlb_apply_batched_f64(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                          int64_t strideA, const double *tau, int64_t strideTau, double *B, int64_t ldb,
                          int64_t strideB, int64_t batch, int32_t *info_out) {
    if (mode == 0) return qlb::dispatch_apply_n<double, 0>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    if (mode == 1) return qlb::dispatch_apply_n<double, 1>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    return qlb::dispatch_apply_n<double, 2>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
}
int qlb_apply_batched_f32(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const float *A, int64_t lda,
                          int64_t strideA, const float *tau, int64_t strideTau, float *B, int64_t ldb,
                          int64_t strideB, int64_t batch, int32_t *info_out) {
    if (mode == 0) return qlb::dispatch_apply_n<float, 0>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    if (mode == 1) return qlb::dispatch_apply_n<float, 1>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    return qlb::dispatch_apply_n<float, 2>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
}
