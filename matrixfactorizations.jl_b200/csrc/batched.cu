// batched.cu -- K6: batched small square Householder QL, Q-apply and square solve (sm_100a).
//
// Per-matrix semantics = the reference's ql! for BlasFloat (LAPACK ?geqlf -> ?geql2 + ?larfg sign
// conventions, reference src/ql.jl:72 with the unblocked spec at src/ql.jl:56-68), lmul!(Q,B) /
// lmul!(Q',B) (src/ql.jl:321-347 / 350-377) and the square ldiv!(F::QL,B) (src/ql.jl:440-447,467).
//
// Factorization kernel (n <= 32): a group of LPM = N/CPL lanes owns one matrix; every lane keeps CPL whole columns in registers
// (a SHIFTED tile: the pivot row always sits at register index N-1, finished rows leave into the shared-memory image of the
// output), so the reflector dot products v^T a_j and the rank-1 update are lane-local FMAs (no shuffles); only the pivot
// column is broadcast (through shared memory) and two scalars are shuffled per step.  A matrix is staged global -> shared by
// ONE TMA tensor copy (3-D tensor map {rows, cols, matrix}, box {N+E, N, 1}: the E out-of-bounds rows give the padded,
// conflict-free image for free) and leaves the same way (one TMA tensor store of the image).  HBM traffic = read A + write
// factors + write tau = (2 n^2 + n) * sizeof(T) bytes per matrix, the compulsory minimum (n = 32 Float64 above 48k matrices
// runs as two kernels and re-reads/re-writes the live 16 x 16 block: 1.22x).  Orders other than 8/16/32 use the next larger
// tile with the matrix in its bottom-right corner; 32 < n <= 64: batched_wy.cuh.
#include "common.cuh"
#include "internal.h"
#include <cuda.h>      // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace qlb {

template <typename T> struct Vec16;
template <> struct Vec16<double> { using type = double2; static constexpr int E = 2; };
template <> struct Vec16<float>  { using type = float4;  static constexpr int E = 4; };

template <typename T> struct Lim;
template <> struct Lim<double> { static constexpr double LO = 1e-280, HI = 1e280; };
template <> struct Lim<float>  { static constexpr float  LO = 1e-24f, HI = 1e30f; };
// ssq >= LO && tot <= HI for non-negative ssq, tot (NaN fails), as integer compares of the high words: the two DSETP
// of the plain form sit on the FP64 pipe, which is the scarce unit of these kernels.
__device__ __forceinline__ bool in_plain_range(double ssq, double tot) {
    // 1e-280 = 0x05CD0B15'A491EB84 -> high word rounded UP (conservative); 1e280 = 0x7A11A0FC'668AAC70 -> rounded DOWN
    return ((unsigned)__double2hiint(ssq) - 0x05CD0B16u <= 0x7FEFFFFFu - 0x05CD0B16u) && ((unsigned)__double2hiint(tot) <= 0x7A11A0FCu);
}
__device__ __forceinline__ bool in_plain_range(float ssq, float tot) { return (ssq >= Lim<float>::LO) && (tot <= Lim<float>::HI); }

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


// ---- TMA tensor copies (SASS: UTMALDG / UTMASTG) ----------------------------------------------------
// The batch is described to the TMA unit as a 3-D tensor {rows = N, cols = N, matrix}; the box is
// {N + E, N, 1}: the E rows past the end of each column are out of bounds, so loads fill them with zeros
// and stores skip them -- one copy per matrix produces / consumes the PADDED shared-memory image
// (column pitch N + E, conflict-free for "one column per lane" accesses).
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int c0, int c1, int c2, const void *smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1),
                 "r"(c2), "r"(smem_u32(smem_src))
                 : "memory");
}

// 32-bit shared-window addressing (keeps the per-step address arithmetic to one add)
__device__ __forceinline__ double2 lds16_s(uint32_t a, double) {
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ float4 lds16_s(uint32_t a, float) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts16_s(uint32_t a, double2 v) { asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void sts16_s(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_s(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_s(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

template <typename T, int N, int CPL>
struct BatchedCfg {
    static constexpr int E = Vec16<T>::E;            // elements per 16-byte vector
    static constexpr int LPM = N / CPL;              // lanes per matrix
    static constexpr int G = 32 / LPM;               // matrices per warp
    static constexpr int PITCH = N + E;              // padded column pitch (elements) -> no bank conflicts
    static constexpr int NCOLS = G * N;              // columns staged per warp-batch
    static constexpr int XPITCH = N + E;             // per-group pivot-column broadcast buffer
    static constexpr int BUF_BYTES = NCOLS * PITCH * (int)sizeof(T);
    static constexpr int X_BYTES = 2 * G * XPITCH * (int)sizeof(T);      // double-buffered
    static constexpr int IMG_BYTES = N * PITCH * (int)sizeof(T);   // one matrix image (multiple of 128 bytes)
    static constexpr int WARP_BYTES = (BUF_BYTES + X_BYTES + 16 + 127) / 128 * 128;   // + mbarrier slot, 128-byte granular
};

// 1/sqrt(x) and 1/x for x in the normal range: hardware seed (MUFU.RSQ64H / RCP64H, ~2^-22) plus ONE
// third-order correction in FMA arithmetic (error ~ seed^3, i.e. below half an ulp before the final
// rounding) -- no special-case branches (callers guarantee the range).
__device__ __forceinline__ double nr_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);               // 1 - x y^2
    const double p = fma(0.375, e, 0.5);
    const double q = y * e;
    return fma(q, p, y);                            // y (1 + e/2 + 3 e^2/8)
}
__device__ __forceinline__ float nr_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double nr_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double p = fma(e, e, e);
    return fma(y, p, y);                            // y (1 + e + e^2)
}
__device__ __forceinline__ float nr_rcp(float x) { return __frcp_rn(x); }

// ---------------------------------------------------------------------------------------------
// Register-tile factorization, organised as LOOPS so that the hot code stays inside the 32 KB
// instruction cache (tools/ubench/icache.cu: straight-line code beyond ~32 KB is fetch-bound at
// ~0.3 instructions/cycle per SM sub-partition -- the fully unrolled variant of this kernel ran at
// exactly that rate no matter how many warps were resident).
//
// The tile is kept SHIFTED: after s steps register index i holds matrix row i-s, so the pivot row of
// step s (R = N-1-s) always sits at index N-1 and every register index is a compile-time constant.
// The shift is free -- the rank-1 update writes a[i+1] = fma(coef, x_i, a[i]).  The finished row R
// leaves the tile each step (straight into the shared-memory image of the output, reflector entries
// already scaled) and a zero is shifted in at the bottom, so finished positions contribute exact
// zeros to the dot products and need no masks.  Steps are grouped (4 or 8 at a time): a group has a
// static lower index LO (indices below hold only zeros) and a static number of live column slots.
// ---------------------------------------------------------------------------------------------

// The Householder steps of the group that starts after LO steps.  NS = column slots still in registers (slot c holds
// column jl + c*LPM); the pivot columns of this group all live in slot NS-1.
template <typename T, int N, int CPL, int LO, int NS, int PD>
__device__ __forceinline__ void ql_group(T (&a)[CPL][N], T (&myscale)[CPL], T *xg, T *obuf, int jl, bool &bad, int nsteps) {
    using V = typename Vec16<T>::type;
    using Cfg = BatchedCfg<T, N, CPL>;
    constexpr int E = Cfg::E, LPM = Cfg::LPM, PITCH = Cfg::PITCH, XP = Cfg::XPITCH;
    constexpr int OWN_SLOT = NS - 1;
    constexpr int V0 = LO / E, V1 = N / E;           // 16-byte vectors covering indices [LO, N)
    constexpr int NVEC = V1 - V0;
    constexpr int NCH = (NS >= 2) ? 2 : 4;           // accumulation chains per slot (fixed summation order)
    constexpr int P = N - 1;                         // register index of the pivot row
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(LO % E == 0 && N % E == 0, "vector alignment");
    const uint32_t xg_s = smem_u32(xg);
    uint32_t stash_s[NS];                            // shared address of this lane's column of slot c in the image
#pragma unroll
    for (int c = 0; c < NS; ++c) stash_s[c] = smem_u32(obuf + (jl + c * LPM) * PITCH);
    // 1. the owner of the group's first pivot column publishes it (indices [LO, N): zeros, then the tail x, then the pivot).
    //    Inside the group the NEXT owner publishes from within the update pass (predicated stores, no divergent branch on
    //    the step's critical path: the branch + reconvergence cost ~150 cycles of every step, ncu source page).
    if (jl == (N - 1 - LO) - OWN_SLOT * LPM) {
        const uint32_t xa0 = xg_s + (uint32_t)((LO & 1) * XP * (int)sizeof(T));
#pragma unroll
        for (int v = V0; v < V1; ++v) {
            V q;
            T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
            for (int e = 0; e < E; ++e) qp[e] = a[OWN_SLOT][v * E + e];
            sts16_s(xa0 + (uint32_t)(v * 16), q);
        }
    }
#pragma unroll 1
    for (int t = 0; t < nsteps; ++t) {
        const int s = LO + t, R = N - 1 - s;
        const int own_lane = R - OWN_SLOT * LPM;
        const uint32_t xa = xg_s + (uint32_t)((s & 1) * XP * (int)sizeof(T));
        const uint32_t xn = xg_s + (uint32_t)(((s + 1) & 1) * XP * (int)sizeof(T));      // next step's buffer
        const bool pub = (jl == own_lane - 1) && (t + 1 < nsteps);                         // this lane owns the next pivot column
        __syncwarp();
        // 2. d_c = sum_{i < P} x_i a[c][i]  (the owner's own column gives the tail's sum of squares)
        T d[NS][NCH];
#pragma unroll
        for (int c = 0; c < NS; ++c)
#pragma unroll
            for (int k = 0; k < NCH; ++k) d[c][k] = T(0);
        // (x is read through an explicit queue of PD vectors: shared-memory latency under load is ~60-80
        //  cycles, one vector feeds only ~9 cycles of FP64 issue, and ptxas by itself keeps just 2 in flight)
        V qa[NVEC];
#pragma unroll
        for (int w = 0; w < PD && w < NVEC; ++w) qa[w] = lds16_s(xa + (uint32_t)((V0 + w) * 16), T());
#pragma unroll
        for (int w = 0; w < NVEC; ++w) {
            if (w + PD < NVEC) qa[w + PD] = lds16_s(xa + (uint32_t)((V0 + w + PD) * 16), T());
            const T *qp = reinterpret_cast<const T *>(&qa[w]);
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int i = (V0 + w) * E + e;
                if (i < P) {
#pragma unroll
                    for (int c = 0; c < NS; ++c) d[c][i % NCH] = t_fma(qp[e], a[c][i], d[c][i % NCH]);
                }
            }
        }
        // the update pass walks x downwards: its first PD vectors are the last ones of the dot pass and stay in registers
        // (the shared-memory data pipe is the busiest unit of this kernel: 64 % of peak wavefronts, ncu)
        // (keeping more than PD of them -- all of x for the short tiles -- was measured 1-5 % slower: registers live across the scalar chain)
        constexpr int RU = (PD < NVEC) ? PD : NVEC;
        V qb[NVEC];
#pragma unroll
        for (int w = 0; w < RU; ++w) qb[w] = qa[NVEC - 1 - w];
        T ds[NS];
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            if constexpr (NCH == 2) ds[c] = d[c][0] + d[c][1];
            else ds[c] = (d[c][0] + d[c][1]) + (d[c][2] + d[c][3]);
        }
        const T ssq = __shfl_sync(FULL, ds[OWN_SLOT], own_lane, LPM);
        const T alpha = __shfl_sync(FULL, a[OWN_SLOT][P], own_lane, LPM);
        const T tot = t_fma(alpha, alpha, ssq);
        // ?larfg's special exits (zero tail -> tau = 0; under/overflowing squares -> scaled norm and the
        // safmin loop) are not taken here: the matrix is flagged and redone by geql2_generic_kernel
        bad = bad || !in_plain_range(ssq, tot);
        // 3. beta = -sign(alpha) ||[alpha;x]||, tau = (beta-alpha)/beta = 1 + |alpha|/||.||, scale = 1/(alpha-beta)
        const T rinv = nr_rsqrt(tot);
        const T nrm = tot * rinv;
        const T beta = -t_copysign(nrm, alpha);
        const T tauk = t_fma(t_abs(alpha), rinv, T(1));
        const T scale = nr_rcp(alpha - beta);
        const T g = -tauk * scale;
        if (jl == own_lane) {
            myscale[OWN_SLOT] = scale;                        // v_i = x_i / (alpha - beta), applied when row i leaves
            sts_s(stash_s[OWN_SLOT] + (uint32_t)(N * (int)sizeof(T)), tauk);      // tau_R waits in the padding row of its column
        }
        // 4. pivot row: finish it, send it to the output image, and the multipliers of the rank-1 update
        T dc[NS];
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            const int col = jl + c * LPM;
            const T aR = a[c][P];
            T sc = t_fma(ds[c], scale, aR);
            sc = (col < R) ? sc : T(0);                       // only columns left of the pivot are updated
            const T newR = t_fma(-tauk, sc, aR);
            dc[c] = g * sc;                                   // multiplier on the raw tail x
            const T val = (col > R) ? aR * myscale[c] : ((col == R) ? beta : newR);
            sts_s(stash_s[c] + (uint32_t)(R * (int)sizeof(T)), val);
        }
        // 5. a_j -= tau (v^T a_j) v on the tail rows, shifting the tile up by one index
#pragma unroll
        for (int w = 0; w < NVEC; ++w) {
            if (w + PD < NVEC && w + PD >= RU) qb[w + PD] = lds16_s(xa + (uint32_t)((V1 - 1 - w - PD) * 16), T());
            const T *qp = reinterpret_cast<const T *>(&qb[w]);
#pragma unroll
            for (int e = E - 1; e >= 0; --e) {
                const int i = (V1 - 1 - w) * E + e;
                if (i < P) {
#pragma unroll
                    for (int c = 0; c < NS; ++c) a[c][i + 1] = t_fma(dc[c], qp[e], a[c][i]);
                }
            }
            // vector (V1 - w) of the shifted tile is final now (its lowest index was written by this w, the others by w - 1):
            // the next owner publishes it
            if (w >= 1 && pub) {
                V q;
                T *qq = reinterpret_cast<T *>(&q);
#pragma unroll
                for (int e = 0; e < E; ++e) qq[e] = a[OWN_SLOT][(V1 - w) * E + e];
                sts16_s(xn + (uint32_t)((V1 - w) * 16), q);
            }
        }
#pragma unroll
        for (int c = 0; c < NS; ++c) a[c][LO] = T(0);
        if (pub) {      // the lowest vector: index LO is the zero that was just shifted in
            V q;
            T *qq = reinterpret_cast<T *>(&q);
#pragma unroll
            for (int e = 0; e < E; ++e) qq[e] = a[OWN_SLOT][V0 * E + e];
            sts16_s(xn + (uint32_t)(V0 * 16), q);
        }
    }
}

// Slot C has no live column any more (all of them are right of every remaining pivot): its registers
// hold reflector entries of rows [0, C*LPM) at indices [N - C*LPM, N); scale them and send them out.
template <typename T, int N, int CPL, int C>
__device__ __forceinline__ void ql_flush_slot(T (&a)[CPL][N], T (&myscale)[CPL], T *obuf, int jl) {
    using V = typename Vec16<T>::type;
    using Cfg = BatchedCfg<T, N, CPL>;
    constexpr int E = Cfg::E, LPM = Cfg::LPM, PITCH = Cfg::PITCH;
    constexpr int S = N - C * LPM;
    T *dst = obuf + (jl + C * LPM) * PITCH;
#pragma unroll
    for (int i = S; i < N; i += E) {
        V q;
        T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
        for (int e = 0; e < E; ++e) qp[e] = a[C][i + e] * myscale[C];
        *reinterpret_cast<V *>(dst + (i - S)) = q;
    }
}

// Group sizes: 4 steps while more than 16 rows are live, 8 afterwards (short tiles: the zero rows cost
// little, and fewer loop bodies keep the code small).  LO is the number of steps already done.
template <int N, int LO, int LPM>
struct GroupSize { static constexpr int value = ((N - LO > 16) ? 4 : 8) < LPM ? ((N - LO > 16) ? 4 : 8) : LPM; };

// STOP > 0: leave after STOP steps (phase A of the split factorization): the live leading block (rows and columns
// below N - STOP, held by slot 0 at register indices [STOP, N)) goes back to the image unscaled and is factored by the
// kernel of the smaller size afterwards.
template <typename T, int N, int CPL, int LO, int PD, int STOP = 0>
struct GroupLoop {
    // ns = number of Householder steps = order of the matrix (ns <= N): a smaller matrix sits in the bottom-right corner of
    // the tile (rows and columns [N - ns, N)), the rest of the tile is zero and its steps are simply not taken.
    static __device__ __forceinline__ void run(T (&a)[CPL][N], T (&myscale)[CPL], T *xg, T *obuf, int jl, bool &bad, int ns) {
        if constexpr (STOP > 0 && LO == STOP) {
            using V = typename Vec16<T>::type;
            constexpr int E = Vec16<T>::E;
            static_assert((N - STOP) * CPL == N || CPL == 1, "the leading block is exactly slot 0");
            T *dst = obuf + jl * BatchedCfg<T, N, CPL>::PITCH;
#pragma unroll
            for (int i = STOP; i < N; i += E) {
                V q;
                T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
                for (int e = 0; e < E; ++e) qp[e] = a[0][i + e];
                *reinterpret_cast<V *>(dst + (i - STOP)) = q;
            }
            return;
        } else {
        constexpr int LPM = N / CPL, GS = GroupSize<N, LO, LPM>::value;
        constexpr int NS = (N - 1 - LO) / LPM + 1;
        constexpr bool LAST = (LO + GS == N);
        static_assert((N - 1 - LO) / LPM == (N - LO - GS) / LPM, "a group's pivot columns live in one slot");
        // steps LO .. ns-2 are full Householder steps; step ns-1 (the matrix's first column: a length-1 reflector is the
        // identity, tau = 0, beta = alpha) only sends the pivot row out
        int nfull = ns - 1 - LO;
        nfull = nfull > GS ? GS : nfull;
        if (nfull > 0) ql_group<T, N, CPL, LO, NS, PD>(a, myscale, xg, obuf, jl, bad, nfull);
        if (LAST || ns - 1 < LO + GS) {
            const int R = N - ns;                       // pivot row of the last step, at register index N-1
#pragma unroll
            for (int c = 0; c < NS; ++c) {
                const int col = jl + c * LPM;
                const T aR = a[c][N - 1];
                obuf[col * BatchedCfg<T, N, CPL>::PITCH + R] = (col > R) ? aR * myscale[c] : aR;
            }
            if (jl == R % LPM) obuf[R * BatchedCfg<T, N, CPL>::PITCH + N] = T(0);
            return;
        }
        if constexpr (!LAST) {
            constexpr int NS_NEXT = (N - 1 - LO - GS) / LPM + 1;
            if constexpr (NS_NEXT < NS) ql_flush_slot<T, N, CPL, NS - 1>(a, myscale, obuf, jl);
            GroupLoop<T, N, CPL, LO + GS, PD, STOP>::run(a, myscale, xg, obuf, jl, bad, ns);
        }
        }
    }
};

template <typename T>
__device__ __noinline__ void geql2_warp(T *__restrict__ Am, int64_t lda, T *__restrict__ taum, int n, T *S, int lane);

// One warp factors G = 32/LPM matrices at a time: TMA (or plain loads) -> shared-memory image ->
// registers -> factor (finished rows go back into the image) -> TMA bulk stores (or plain stores).
// A matrix whose squares leave the plain range (flagged, rare) is left untouched in global memory by the register path and
// factored right away by the same warp with the literal ?geql2 / ?larfg routine (geql2_warp) -- one launch per call, no
// redo list, no counter to clear.
// STOP > 0: phase A only (see GroupLoop); flags[mi] (written for EVERY matrix) = 1 marks the matrices that were flagged
// and therefore factored completely here.  skip != null: matrices with skip[mi] != 0 are not touched (phase B).
template <typename T, int N, int CPL, int WARPS, int MINB, bool USE_TMA, int PD = 4, int STOP = 0, bool PADDED = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
geqlf_batched_kernel(const __grid_constant__ CUtensorMap tmap, T *__restrict__ A, int64_t lda, int64_t strideA,
                     T *__restrict__ tau, int64_t strideTau, int64_t batch, int *__restrict__ flags,
                     const int *__restrict__ skip, int nr_arg) {
    using Cfg = BatchedCfg<T, N, CPL>;
    using V = typename Vec16<T>::type;
    constexpr int E = Cfg::E, LPM = Cfg::LPM, G = Cfg::G, PITCH = Cfg::PITCH;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LPM, jl = lane % LPM;    // (matrices interleaved lane by lane instead: 10-25 % slower, bank conflicts on the images)
    unsigned char *wbase = smem_raw + (size_t)warp * Cfg::WARP_BYTES;
    T *buf = reinterpret_cast<T *>(wbase);
    T *obuf = buf + g * N * PITCH;                           // this matrix's image (input, then output)
    T *xg = reinterpret_cast<T *>(wbase + Cfg::BUF_BYTES) + g * 2 * Cfg::XPITCH;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + Cfg::BUF_BYTES + Cfg::X_BYTES);

    const int nbatch = (int)batch;                           // < 2^31 (checked by the launcher)
    const int nwb = (nbatch + G - 1) / G;                    // warp-batches of G matrices
    const int gw = (int)blockIdx.x * WARPS + warp, nw = (int)gridDim.x * WARPS;
    // order nr <= N: the matrix occupies rows and columns [pad, N) of the tile; the tensor copies start at (-pad, -pad), so the
    // out-of-bounds part of the box is zero-filled on the way in and skipped on the way out
    // (PADDED is a template parameter: the exact-size kernels keep nr = N as a constant -- the extra live registers cost the
    //  n = 32 Float64 kernel 24 bytes of spills)
    const int nr = PADDED ? nr_arg : N;
    const int pad = N - nr;

    if constexpr (USE_TMA) {
        if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
        __syncwarp();
    }
    uint32_t phase = 0;
    for (int wb = gw; wb < nwb; wb += nw) {
        const int m0 = wb * G, mi = m0 + g;
        const bool inb = mi < nbatch;                                  // inside the batch (its image is loaded)
        const bool valid = inb && !(skip != nullptr && skip[mi] != 0);
        // ---- stage the matrices of this warp-batch ----
        if constexpr (USE_TMA) {
            const int nvalid = (nbatch - m0) < G ? (nbatch - m0) : G;
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)(nvalid * Cfg::IMG_BYTES));
                for (int q = 0; q < nvalid; ++q) tma_load_3d(buf + q * N * PITCH, &tmap, -pad, -pad, m0 + q, bar);
            }
            if (!inb) {        // tail of the batch: an identity keeps the arithmetic finite (results are discarded)
                for (int idx = jl; idx < N * N; idx += LPM) obuf[(idx / N) * PITCH + (idx % N)] = (idx / N == idx % N) ? T(1) : T(0);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            __syncwarp();
        } else {
            const T *Am = A + (int64_t)(inb ? mi : 0) * strideA;
            for (int idx = jl; idx < N * N; idx += LPM) {
                const int c = idx / N, r = idx % N;
                T v = (c == r) ? T(1) : T(0);
                if (inb) v = (c >= pad && r >= pad) ? Am[(int64_t)(c - pad) * lda + (r - pad)] : T(0);
                obuf[c * PITCH + r] = v;
            }
            __syncwarp();
        }
        T a[CPL][N];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int col = jl + c * LPM;
            const V *src = reinterpret_cast<const V *>(obuf + col * PITCH);
#pragma unroll
            for (int v = 0; v < N / E; ++v) {
                V q = src[v];
                const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
                for (int e = 0; e < E; ++e) a[c][v * E + e] = qp[e];
            }
        }
        T myscale[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) myscale[c] = T(0);

        bool bad = false;
        GroupLoop<T, N, CPL, 0, PD, STOP>::run(a, myscale, xg, obuf, jl, bad, nr);

        // flagged matrices stay untouched in global memory (no tau, no store) and are redone below
        if (flags != nullptr && valid && jl == 0) flags[mi] = bad ? 1 : 0;
        const unsigned redo = __ballot_sync(0xffffffffu, valid && bad && jl == 0);
        __syncwarp();
        if (valid && !bad) {                                      // tau: one coalesced row per matrix
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const int col = jl + c * LPM;
                if ((STOP == 0 || col >= N - STOP) && col >= pad) tau[(int64_t)mi * strideTau + col - pad] = obuf[col * PITCH + N];      // phase A: the rest comes from phase B
            }
        }
        // (tensor STORES that start at negative coordinates fault on sm_100 -- tools/ubench/tma_oob.cu: loads zero-fill, stores
        //  raise "illegal instruction" -- so the padded variant sends its corner out with plain stores)
        if constexpr (USE_TMA && !PADDED) {
            fence_proxy_async();                                  // generic-proxy writes -> visible to the bulk copies
            __syncwarp();
            if (valid && !bad && jl == 0) tma_store_3d(&tmap, 0, 0, mi, obuf);
            tma_store_commit();
            tma_store_wait_read();                                // the image may be overwritten by the next load
            __syncwarp();
        } else {
            if (valid && !bad) {
                T *Am = A + (int64_t)mi * strideA;
                if (PADDED && USE_TMA && (pad % E) == 0) {       // (USE_TMA: base, lda and strideA are 16-byte aligned) whole vectors
                    const int nv = nr / E;
                    for (int c = 0; c < nr; ++c)
                        for (int v = jl; v < nv; v += LPM)
                            *reinterpret_cast<V *>(Am + (int64_t)c * lda + v * E) = *reinterpret_cast<const V *>(obuf + (c + pad) * PITCH + pad + v * E);
                } else {
                    for (int c = 0; c < nr; ++c)
                        for (int r = jl; r < nr; r += LPM) Am[(int64_t)c * lda + r] = obuf[(c + pad) * PITCH + pad + r];
                }
            }
            if constexpr (USE_TMA) fence_proxy_async();      // generic reads of the image before the next bulk load into it
            __syncwarp();
        }
        if (redo) {      // warp-uniform and rare: the images have been read by the stores above, the buffer is free scratch
            static_assert(Cfg::BUF_BYTES >= ((N + 1) * N + N) * (int)sizeof(T), "the warp's image area holds the generic routine's scratch");
            for (unsigned rm = redo; rm; rm &= rm - 1) {
                const int gi = (__ffs(rm) - 1) / LPM;
                geql2_warp<T>(A + (int64_t)(m0 + gi) * strideA, lda, tau + (int64_t)(m0 + gi) * strideTau, nr, buf, lane);
            }
            if constexpr (USE_TMA) fence_proxy_async();      // generic-proxy writes to the buffer before the next bulk load into it
            __syncwarp();
        }
    }
    if constexpr (USE_TMA) tma_store_wait_all();
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

// One warp factors one n x n matrix (column-major, leading dimension lda) in place, through the shared-memory scratch S
// ((n + 1) * n + n elements): LAPACK ?geql2 / ?larfg literally.
template <typename T>
__device__ __noinline__ void geql2_warp(T *__restrict__ Am, int64_t lda, T *__restrict__ taum, int n, T *S, int lane) {
    const int P = n + 1;                                     // column pitch
    T *taus = S + P * n;
    const T safmin = LarfgConst<T>::safmin(), rsafmn = T(1) / safmin;
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
    for (int i = lane; i < n; i += 32) taum[i] = taus[i];
    __syncwarp();
}

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) geql2_generic_kernel(T *__restrict__ A, int64_t lda, int64_t strideA, T *__restrict__ tau,
                                                                  int64_t strideTau, int n, int64_t count) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T *S = reinterpret_cast<T *>(smem_raw) + (size_t)warp * ((n + 1) * n + n);
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
    for (int64_t mi = gw; mi < count; mi += nw) geql2_warp<T>(A + mi * strideA, lda, tau + mi * strideTau, n, S, lane);
}

// Dynamic shared memory attribute + resident blocks per SM of a batched kernel.  Both cost several microseconds of host
// time per query, which is visible at the shard sizes of an 8-GPU run (a 32768-matrix shard takes 0.2 ms): they are
// looked up once per handle (= per device), kernel and size.
template <typename K>
static int kernel_ready(qlb200_ctx *h, K kern, int threads, size_t smem, int *occ) {
    const void *fn = (const void *)kern;
    size_t granted = 0;                    // the attribute is per function: it only ever grows (the generic kernels' size varies with n)
    for (int i = 0; i < h->n_ksetup; ++i)
        if (h->ksetup[i].fn == fn) {
            if (h->ksetup[i].smem == smem) {
                *occ = h->ksetup[i].occ;
                return 0;
            }
            if (h->ksetup[i].smem > granted) granted = h->ksetup[i].smem;
        }
    if (smem > granted)
        QLB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QLB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, threads, smem));
    if (h->n_ksetup < 256) {
        h->ksetup[h->n_ksetup].fn = fn;
        h->ksetup[h->n_ksetup].smem = smem;
        h->ksetup[h->n_ksetup].occ = *occ;
        h->n_ksetup++;
    }       // a full table only costs the lookup again next time: an unrecorded size above every recorded one re-sets the attribute
    return 0;
}

template <typename T>
static int launch_geql2_generic(qlb200_ctx *h, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau, int n, int64_t count) {
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * ((size_t)(n + 1) * n + n) * sizeof(T);
    auto kern = geql2_generic_kernel<T, WARPS>;
    int occ_unused = 0;
    int rcs = kernel_ready(h, kern, WARPS * 32, smem, &occ_unused);
    if (rcs) return rcs;
    int64_t blocks = (count + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)h->sm_count * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(A, lda, strideA, tau, strideTau, n, count);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Q-apply and square solve.  A group of LPM = N*NR/8 lanes owns one matrix (G = 32/LPM matrices per warp:
// 1 for n = 32, 2 for n = 16, 4 for n = 8 when NR = 8 columns of B are processed at a time); lane
// (jb, rg) = (l % NR, l / NR) of a group owns RPL = 8 consecutive rows of column jb of B in registers.
// The factors of the G matrices are staged by one TMA tensor copy each (padded image, as in the
// factorization kernel), double-buffered when the images are small; reflector k's vector is read from the
// image with 16-byte loads, the dot product is finished with log2(RG) xor-shuffles.
// MODE 0: B <- Q B (k ascending, src/ql.jl:321-347); MODE 1: B <- Q' B (k descending,
// src/ql.jl:350-377); MODE 2: B <- L^{-1} Q' B (src/ql.jl:440-447,467).
// ---------------------------------------------------------------------------------------------
template <typename T, int N, int NR, int MODE = 0>
struct ApplyCfg {
    static constexpr int E = Vec16<T>::E;
    static constexpr int PITCH = N + E;
    static constexpr int IMG_BYTES = N * PITCH * (int)sizeof(T);                 // multiple of 128
    static constexpr int pow2_floor(int x) { int p = 1; while (2 * p <= x) p *= 2; return p; }
#ifndef APPLY_CL
#define APPLY_CL 2
#endif
    // columns of B per lane: the kernel is bound by shared-memory wavefronts (every LDS.128 of a reflector is 4 of them, and
    // a quarter warp reads the same 16 bytes), so one reflector load feeds CL columns
    static constexpr int CL = (NR >= 2 && N >= 16 && !(MODE == 2 && N == 16 && sizeof(T) == 4)) ? APPLY_CL : 1;   // (that one: 6 % slower)
    static constexpr int NRL = NR / CL;                                          // lanes across the NR columns of a pass
    static constexpr int LPM0 = (N * NRL / 8) > 32 ? 32 : (N * NRL / 8);         // lanes per matrix at 8 rows per lane
    static constexpr int GMAX = pow2_floor((24 * 1024) / IMG_BYTES > 0 ? (24 * 1024) / IMG_BYTES : 1);   // shared-memory budget per warp
    static constexpr int G = (32 / LPM0) < GMAX ? (32 / LPM0) : GMAX;            // matrices per warp
    static constexpr int LPM = 32 / G;                    // lanes per matrix
    static constexpr int RG = LPM / NRL;                  // row groups per matrix
    static constexpr int RPL = N / RG;                    // rows per lane
    static constexpr int TAU_BYTES = (G * N * (int)sizeof(T) + 127) / 128 * 128;
    static constexpr int STAGE_BYTES = G * IMG_BYTES + TAU_BYTES;
    static constexpr int DBUF = (STAGE_BYTES <= 6 * 1024) ? 2 : 1;              // prefetch the next warp-batch if it is cheap
    // explicit reflectors (unit pivot, zeros below): made in place for the Q-apply, in a second image for the solve (which
    // still needs L); not for n = 8, where every lane holds whole columns and the masks are cheap (measured)
    static constexpr bool CLEAN = (MODE != 2) && (N >= 16);
    static constexpr bool VCOPY = (MODE == 2) && (N >= 16) && (sizeof(T) == 4);     // Float64: the extra shared memory costs more than the masks (measured)
    static constexpr int VIMG_OFF = DBUF * STAGE_BYTES + 128;                     // after the mbarriers
    static constexpr int WARP_BYTES = VIMG_OFF + (VCOPY ? G * IMG_BYTES : 0);
    static constexpr bool VALID = RG >= 1 && RPL % E == 0;      // row blocks are whole 16-byte vectors (else: the next wider NR)
};

template <typename T, int N, int NR, int MODE, int WARPS, bool USE_TMA, bool PADDED = false>
__global__ void __launch_bounds__(WARPS * 32)
apply_batched_kernel(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ A, int64_t lda, int64_t strideA,
                     const T *__restrict__ tau, int64_t strideTau, T *__restrict__ B, int64_t ldb, int64_t strideB, int nrhs,
                     int64_t batch, int32_t *__restrict__ info_out, int vecB, int nr_arg) {
    // PADDED: factors of order nr < N in the bottom-right corner of the N x N tile (zero rows/columns and tau = 0 in front:
    // those reflectors are the identity), B in rows [pad, N) -- same arithmetic on the real part, zeros elsewhere.
    const int nr = PADDED ? nr_arg : N;
    const int pad = N - nr;
    using Cfg = ApplyCfg<T, N, NR, MODE>;
    using V = typename Vec16<T>::type;
    constexpr int E = Cfg::E, LPM = Cfg::LPM, G = Cfg::G, RPL = Cfg::RPL, PITCH = Cfg::PITCH, DBUF = Cfg::DBUF;
    constexpr int CL = Cfg::CL, NRL = Cfg::NRL;
    constexpr bool CLEAN = Cfg::CLEAN, VCOPY = Cfg::VCOPY;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LPM, l = lane % LPM;
    const int jb = l % NRL, rg = l / NRL, row0 = rg * RPL;      // lane (jb, rg): columns jb*CL .. jb*CL+CL-1 of the pass, rows row0 ..
    unsigned char *wbase = smem_raw + (size_t)warp * Cfg::WARP_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + DBUF * Cfg::STAGE_BYTES);
    const int nbatch = (int)batch;
    const int nwb = (nbatch + G - 1) / G;
    const int gw = (int)blockIdx.x * WARPS + warp, nw = (int)gridDim.x * WARPS;
    if constexpr (USE_TMA) {
        if (lane == 0) {
            for (int s = 0; s < DBUF; ++s) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    // stage the factors and tau of warp-batch wb into stage s
    auto issue = [&](int wb, int s) {
        unsigned char *st = wbase + s * Cfg::STAGE_BYTES;
        T *img = reinterpret_cast<T *>(st);
        T *taus = reinterpret_cast<T *>(st + G * Cfg::IMG_BYTES);
        const int m0 = wb * G;
        const int nvalid = (nbatch - m0) < G ? (nbatch - m0) : G;
        if constexpr (USE_TMA) {
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[s], (uint32_t)(nvalid * (Cfg::IMG_BYTES + nr * (int)sizeof(T))));
                for (int q = 0; q < nvalid; ++q) {
                    tma_load_3d(img + q * N * PITCH, &tmap, -pad, -pad, m0 + q, &bars[s]);
                    tma_load_1d(taus + q * N + pad, tau + (int64_t)(m0 + q) * strideTau, nr * sizeof(T), &bars[s]);
                }
            }
        } else {
            for (int q = 0; q < nvalid; ++q) {
                const T *Am = A + (int64_t)(m0 + q) * strideA;
                for (int idx = lane; idx < N * N; idx += 32) {
                    const int c = idx / N, r = idx % N;
                    img[q * N * PITCH + c * PITCH + r] = (c >= pad && r >= pad) ? Am[(int64_t)(c - pad) * lda + (r - pad)] : T(0);
                }
                for (int i = lane; i < N; i += 32) taus[q * N + i] = (i >= pad) ? tau[(int64_t)(m0 + q) * strideTau + i - pad] : T(0);
            }
        }
    };
    uint32_t phase[DBUF];
#pragma unroll
    for (int s = 0; s < DBUF; ++s) phase[s] = 0;
    if constexpr (PADDED && USE_TMA) {      // tau of the padded reflectors: zero, never overwritten by the bulk copies
        for (int s = 0; s < DBUF; ++s) {
            T *taus = reinterpret_cast<T *>(wbase + s * Cfg::STAGE_BYTES + G * Cfg::IMG_BYTES);
            for (int i = lane; i < G * N; i += 32)
                if (i % N < pad) taus[i] = T(0);
        }
        fence_proxy_async();
        __syncwarp();
    }
    if constexpr (DBUF == 2) {
        if (gw < nwb) issue(gw, 0);
    }
    int it = 0;
    for (int wb = gw; wb < nwb; wb += nw, ++it) {
        const int cur = (DBUF == 2) ? (it & 1) : 0;
        if constexpr (DBUF == 2) {
            if (wb + nw < nwb) issue(wb + nw, cur ^ 1);         // that stage was released at the end of the last iteration
        } else {
            issue(wb, 0);
        }
        const int mi = wb * G + g;
        const bool valid = mi < nbatch;
        unsigned char *st = wbase + cur * Cfg::STAGE_BYTES;
        const T *img = reinterpret_cast<const T *>(st) + g * N * PITCH;
        const T *taus = reinterpret_cast<const T *>(st + G * Cfg::IMG_BYTES) + g * N;
        bool waited = false;
        // this lane's slice of B: nrhs columns are processed NR at a time
        for (int c0 = 0; c0 < nrhs; c0 += NR) {
            bool cvalid[CL];
            T b[CL][RPL];
            T *Bc[CL];
#pragma unroll
            for (int cc = 0; cc < CL; ++cc) {
                const int col = c0 + jb * CL + cc;
                cvalid[cc] = valid && col < nrhs;
                Bc[cc] = B + (int64_t)(valid ? mi : 0) * strideB + (int64_t)(col < nrhs ? col : 0) * ldb + row0 - pad;
                if (PADDED) {
#pragma unroll
                    for (int i = 0; i < RPL; ++i) b[cc][i] = (cvalid[cc] && row0 + i >= pad) ? Bc[cc][i] : T(0);
                } else if (vecB) {
#pragma unroll
                    for (int v = 0; v < RPL / E; ++v) {
                        V q = cvalid[cc] ? reinterpret_cast<const V *>(Bc[cc])[v] : V{};
                        const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
                        for (int e = 0; e < E; ++e) b[cc][v * E + e] = qp[e];
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < RPL; ++i) b[cc][i] = cvalid[cc] ? Bc[cc][i] : T(0);
                }
            }
            if (!waited) {
                if constexpr (USE_TMA) {
                    mbar_wait(&bars[cur], phase[cur]);
                    phase[cur] ^= 1;
                }
                __syncwarp();
                waited = true;
                if constexpr (CLEAN) {
                    // Q-apply only: make the staged reflectors explicit once per warp-batch (unit pivot, zeros below it --
                    // L is not needed), so that the reflector loop below carries no row masks.
                    T *imw = reinterpret_cast<T *>(st);
                    for (int col = lane; col < G * N; col += 32) {
                        const int k = col % N;
                        T *cp = imw + (col / N) * N * PITCH + k * PITCH;
                        V q = *reinterpret_cast<const V *>(cp + (k / E) * E);
                        T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
                        for (int e = 0; e < E; ++e) qp[e] = (e < k % E) ? qp[e] : ((e == k % E) ? T(1) : T(0));
                        *reinterpret_cast<V *>(cp + (k / E) * E) = q;
                        for (int v = k / E + 1; v < N / E; ++v) *reinterpret_cast<V *>(cp + v * E) = V{};
                    }
                    __syncwarp();
                }
                if constexpr (VCOPY) {
                    const T *imr = reinterpret_cast<const T *>(st);
                    T *vw = reinterpret_cast<T *>(wbase + Cfg::VIMG_OFF);
                    for (int col = lane; col < G * N; col += 32) {
                        const int k = col % N, off = (col / N) * N * PITCH + k * PITCH;
                        for (int v = 0; v < k / E; ++v) *reinterpret_cast<V *>(vw + off + v * E) = *reinterpret_cast<const V *>(imr + off + v * E);
                        V q = *reinterpret_cast<const V *>(imr + off + (k / E) * E);
                        T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
                        for (int e = 0; e < E; ++e) qp[e] = (e < k % E) ? qp[e] : ((e == k % E) ? T(1) : T(0));
                        *reinterpret_cast<V *>(vw + off + (k / E) * E) = q;
                        for (int v = k / E + 1; v < N / E; ++v) *reinterpret_cast<V *>(vw + off + v * E) = V{};
                    }
                    __syncwarp();
                }
            }
            // reflectors: column k (0-based) has pivot row k, tail rows 0..k-1
#pragma unroll 4
            for (int kk = 0; kk < N; ++kk) {
                const int k = (MODE == 0) ? kk : N - 1 - kk;
                const T *vk = (VCOPY ? reinterpret_cast<const T *>(wbase + Cfg::VIMG_OFF) + g * N * PITCH : img) + k * PITCH + row0;
                const T tk = valid ? taus[k] : T(0);
                const int kd = k - row0;                       // rows with local index < kd belong to the tail
                T v[RPL];
                T part[CL];
#pragma unroll
                for (int cc = 0; cc < CL; ++cc) part[cc] = T(0);
#pragma unroll
                for (int vv = 0; vv < RPL / E; ++vv) {
                    V q = *reinterpret_cast<const V *>(vk + vv * E);
                    const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int i = vv * E + e;
                        if constexpr (!CLEAN && !VCOPY) v[i] = (i < kd) ? qp[e] : (i == kd ? T(1) : T(0));
                        else v[i] = qp[e];
#pragma unroll
                        for (int cc = 0; cc < CL; ++cc) part[cc] = t_fma(v[i], b[cc][i], part[cc]);
                    }
                }
#pragma unroll
                for (int off = NRL; off < LPM; off <<= 1) {
#pragma unroll
                    for (int cc = 0; cc < CL; ++cc) part[cc] += __shfl_xor_sync(FULL, part[cc], off);
                }
#pragma unroll
                for (int cc = 0; cc < CL; ++cc) {
                    const T s = tk * part[cc];
#pragma unroll
                    for (int i = 0; i < RPL; ++i) b[cc][i] = t_fma(-s, v[i], b[cc][i]);
                }
            }
            if constexpr (MODE == 2) {
                // forward substitution with L = tril(factors): x_c = b_c / L_cc; b_i -= L_ic x_c (i > c).  The reciprocals of
                // the diagonal do not depend on the substitution: they are computed up front (into the padding element of
                // each image column), so the serial chain carries a multiplication instead of a division.
                T *imgw = const_cast<T *>(img);
                if (c0 == 0) {
                    for (int c = l; c < N; c += LPM) imgw[c * PITCH + N] = (c >= pad) ? T(1) / img[c * PITCH + c] : T(0);
                    __syncwarp();
                }
                int bad = 0;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const int orow = c % RPL, org = c / RPL;       // owner row group / local row of row c
                    const T *lc = img + c * PITCH + row0;
                    const T rcc = img[c * PITCH + N];
                    if (img[c * PITCH + c] == T(0) && bad == 0 && c >= pad) bad = c + 1 - pad;
                    T xc[CL];
#pragma unroll
                    for (int cc = 0; cc < CL; ++cc) xc[cc] = __shfl_sync(FULL, b[cc][orow], jb + NRL * org, LPM) * rcc;
                    const int cd = c - row0;
#pragma unroll
                    for (int vv = 0; vv < RPL / E; ++vv) {
                        V q = *reinterpret_cast<const V *>(lc + vv * E);
                        const T *qp = reinterpret_cast<const T *>(&q);
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            const int i = vv * E + e;
#pragma unroll
                            for (int cc = 0; cc < CL; ++cc) {
                                if (i == cd) b[cc][i] = xc[cc];
                                else if (i > cd) b[cc][i] = t_fma(-qp[e], xc[cc], b[cc][i]);
                            }
                        }
                    }
                }
                if (info_out != nullptr && valid && l == 0 && c0 == 0) info_out[mi] = bad;
            }
#pragma unroll
            for (int cc = 0; cc < CL; ++cc) {
                if (!cvalid[cc]) continue;
                if (PADDED) {
#pragma unroll
                    for (int i = 0; i < RPL; ++i)
                        if (row0 + i >= pad) Bc[cc][i] = b[cc][i];
                } else if (vecB) {
#pragma unroll
                    for (int v = 0; v < RPL / E; ++v) {
                        V q;
                        T *qp = reinterpret_cast<T *>(&q);
#pragma unroll
                        for (int e = 0; e < E; ++e) qp[e] = b[cc][v * E + e];
                        reinterpret_cast<V *>(Bc[cc])[v] = q;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < RPL; ++i) Bc[cc][i] = b[cc][i];
                }
            }
        }
        if (!waited) {      // nrhs == 0 cannot happen (checked by the caller); keep the barrier phases consistent anyway
            if constexpr (USE_TMA) {
                mbar_wait(&bars[cur], phase[cur]);
                phase[cur] ^= 1;
            }
        }
        if constexpr (USE_TMA) fence_proxy_async();   // this lane's generic-proxy writes to the stage (explicit reflectors, reciprocal diagonal) before the next bulk copy into it
        __syncwarp();    // all lanes are done with this stage before it is refilled
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
// 3-D tensor map {rows N, cols N, matrix} over the strided batch, box {pitch, N, 1} (see tma_load_3d).
// The encoder lives in the driver library; it is fetched through the runtime so that libqlb200.so has no
// link-time dependency on libcuda (the library must load on a box without a GPU driver).
typedef CUresult (*qlb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <typename T>
static bool make_batch_tensor_map(CUtensorMap *tm, const T *A, int n, int64_t lda, int64_t strideA, int64_t batch, int pitch,
                                  int nbox = 0) {
    if (nbox == 0) nbox = n;                // tile order (>= n: the copies then start at negative coordinates, see the kernels)
    static qlb_encode_tiled_fn enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            enc = (qlb_encode_tiled_fn)fn;
        else
            (void)cudaGetLastError();
    }
    if (!enc || pitch > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)batch};
    // a single matrix may come with any stride: use a harmless consistent one
    const int64_t sA = batch > 1 ? strideA : (int64_t)lda * n;
    const cuuint64_t strides[2] = {(cuuint64_t)lda * sizeof(T), (cuuint64_t)sA * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)pitch, (cuuint32_t)nbox, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    // a column shorter than 256 bytes inside a wider leading dimension (phase B of the n = 32 split: 128 of every 256 bytes)
    // must not be promoted to 256-byte fetches: that doubled the DRAM reads of that kernel (ncu: 1.07 GB for 0.54 GB)
    const bool narrow = (size_t)n * sizeof(T) < 256 && lda > n;
    return enc(tm, dt, 3, (void *)A, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               narrow ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
static bool tma_ok(const T *A, int64_t lda, int64_t strideA) {
    return ((uintptr_t)A % 16 == 0) && ((lda * sizeof(T)) % 16 == 0) && ((strideA * sizeof(T)) % 16 == 0);
}

template <typename T, int N, int CPL, int WARPS, int MINB, int PD = 4>
static int launch_geqlf_batched(qlb200_ctx *h, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau,
                                int64_t batch, int nr = N) {
    using Cfg = BatchedCfg<T, N, CPL>;
    const size_t smem = (size_t)WARPS * Cfg::WARP_BYTES;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const bool tma = tma_ok(A, lda, strideA) && make_batch_tensor_map<T>(&tmap, A, nr, lda, strideA, batch, Cfg::PITCH, N);
    auto kern = tma ? geqlf_batched_kernel<T, N, CPL, WARPS, MINB, true, PD> : geqlf_batched_kernel<T, N, CPL, WARPS, MINB, false, PD>;
    if (nr != N) kern = tma ? geqlf_batched_kernel<T, N, CPL, WARPS, MINB, true, PD, 0, true> : geqlf_batched_kernel<T, N, CPL, WARPS, MINB, false, PD, 0, true>;
    int occ = 0;
    int rcs = kernel_ready(h, kern, WARPS * 32, smem, &occ);
    if (rcs) return rcs;
    if (occ < 1) return qlb_fail(h, 1000 + (int)cudaErrorLaunchOutOfResources, "batched QL kernel does not fit on an SM%s");
    const int64_t nwb = (batch + Cfg::G - 1) / Cfg::G;
    int64_t blocks = (nwb + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)h->sm_count * occ;              // persistent: one resident wave
    if (h->opt_bgrid > 0 && cap > h->opt_bgrid) cap = h->opt_bgrid;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (batch >= ((int64_t)1 << 31)) return qlb_fail(h, -8, "batch too large%s");
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(tmap, A, lda, strideA, tau, strideTau, batch, nullptr, nullptr, nr);      // ONE launch
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// n = 32 in two kernels: phase A (R = 31 ... 16) with the 32 x 32 register tile, then the live 16 x 16 leading block is
// an ordinary 16 x 16 QL (leading dimension lda) for the n = 16 kernel, which packs FOUR matrices into a warp instead of
// two: the per-step fixed cost (broadcast round trip, ?larfg chain) of the 16 cheap steps is shared by twice as many
// matrices.  Costs 4 KB of extra HBM traffic per matrix (the block is written and read once more).
template <typename T, int WARPS, int MINB_A, int MINB_B>
static int launch_geqlf_batched_split32(qlb200_ctx *h, T *A, int64_t lda, int64_t strideA, T *tau, int64_t strideTau, int64_t batch) {
    using CfgA = BatchedCfg<T, 32, 2>;
    using CfgB = BatchedCfg<T, 16, 2>;
    if (batch >= ((int64_t)1 << 31)) return qlb_fail(h, -8, "batch too large%s");
    CUtensorMap tmA, tmB;
    memset(&tmA, 0, sizeof(tmA));
    memset(&tmB, 0, sizeof(tmB));
    const bool tma = tma_ok(A, lda, strideA) && make_batch_tensor_map<T>(&tmA, A, 32, lda, strideA, batch, CfgA::PITCH) &&
                     make_batch_tensor_map<T>(&tmB, A, 16, lda, strideA, batch, CfgB::PITCH);
    auto kA = tma ? geqlf_batched_kernel<T, 32, 2, WARPS, MINB_A, true, 4, 16> : geqlf_batched_kernel<T, 32, 2, WARPS, MINB_A, false, 4, 16>;
    auto kB = tma ? geqlf_batched_kernel<T, 16, 2, WARPS, MINB_B, true, 4, 0> : geqlf_batched_kernel<T, 16, 2, WARPS, MINB_B, false, 4, 0>;
    const size_t smA = (size_t)WARPS * CfgA::WARP_BYTES, smB = (size_t)WARPS * CfgB::WARP_BYTES;
    int occA = 0, occB = 0;
    int rcs = kernel_ready(h, kA, WARPS * 32, smA, &occA);
    if (!rcs) rcs = kernel_ready(h, kB, WARPS * 32, smB, &occB);
    if (rcs) return rcs;
    if (occA < 1 || occB < 1) return qlb_fail(h, 1000 + (int)cudaErrorLaunchOutOfResources, "batched QL kernel does not fit on an SM%s");
    auto grid = [&](int G, int occ) {
        const int64_t nwb = (batch + G - 1) / G;
        int64_t blocks = (nwb + WARPS - 1) / WARPS;
        const int64_t cap = (int64_t)h->sm_count * occ;
        if (blocks > cap) blocks = cap;
        return (unsigned)(blocks < 1 ? 1 : blocks);
    };
    // flags[mi]: written by phase A for every matrix (1 = flagged and already factored completely there), read by phase B
    int rc = qlb_reserve(h, &h->redo, &h->redo_bytes, (size_t)batch * sizeof(int));
    if (rc) return rc;
    int *flags = (int *)h->redo;
    kA<<<grid(CfgA::G, occA), WARPS * 32, smA, h->stream>>>(tmA, A, lda, strideA, tau, strideTau, batch, flags, nullptr, 32);
    QLB_LAUNCH_CHECK(h);
    kB<<<grid(CfgB::G, occB), WARPS * 32, smB, h->stream>>>(tmB, A, lda, strideA, tau, strideTau, batch, nullptr, flags, 16);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T, int N, int NR, int MODE>
static int launch_apply_batched(qlb200_ctx *h, const T *A, int64_t lda, int64_t strideA, const T *tau,
                                int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch,
                                int32_t *info_out, int nr = N) {
    using Cfg = ApplyCfg<T, N, NR, MODE>;
    constexpr int WARPS = (N == 64 && 6 * Cfg::WARP_BYTES <= 232448) ? 6 : 4;      // 64-row tiles: one CTA per SM, fill its shared memory
    const size_t smem = (size_t)WARPS * Cfg::WARP_BYTES;
    if (batch >= ((int64_t)1 << 31)) return qlb_fail(h, -8, "batch too large%s");
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const bool tau_ok = ((uintptr_t)tau % 16 == 0) && ((strideTau * sizeof(T)) % 16 == 0 || batch <= 1);
    const bool pad_ok = (nr == N) || (((N - nr) * sizeof(T)) % 16 == 0 && (nr * sizeof(T)) % 16 == 0);      // the 1-D bulk copy of tau
    const bool tma = tma_ok(A, lda, strideA) && tau_ok && pad_ok && make_batch_tensor_map<T>(&tmap, A, nr, lda, strideA, batch, Cfg::PITCH, N);
    const int vecB = ((uintptr_t)B % 16 == 0) && ((ldb * sizeof(T)) % 16 == 0) && ((strideB * sizeof(T)) % 16 == 0 || batch <= 1);
    auto kern = tma ? apply_batched_kernel<T, N, NR, MODE, WARPS, true> : apply_batched_kernel<T, N, NR, MODE, WARPS, false>;
    if (nr != N) kern = tma ? apply_batched_kernel<T, N, NR, MODE, WARPS, true, true> : apply_batched_kernel<T, N, NR, MODE, WARPS, false, true>;
    int occ = 0;
    int rcs = kernel_ready(h, kern, WARPS * 32, smem, &occ);
    if (rcs) return rcs;
    if (occ < 1) return qlb_fail(h, 1000 + (int)cudaErrorLaunchOutOfResources, "batched apply kernel does not fit on an SM%s");
    const int64_t nwb = (batch + Cfg::G - 1) / Cfg::G;
    int64_t blocks = (nwb + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)h->sm_count * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(tmap, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch,
                                                           info_out, vecB, nr);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Generic Q-apply / square solve for any n <= 64 (the sizes the register kernel above is not instantiated for; the
// factorization accepts every n <= 64, so the apply must too).  One warp per matrix; factors (pitch n+1) and tau
// staged in shared memory with plain loads; lane (jb, sub) = (lane % 8, lane / 8) owns rows sub, sub+4, ... of
// column jb of an 8-column chunk of B in registers; a reflector's dot product is finished with two xor-shuffles.
// Same MODE meaning and the same per-matrix arithmetic as apply_batched_kernel.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) apply_generic_kernel(const T *__restrict__ A, int64_t lda, int64_t strideA,
                                                                  const T *__restrict__ tau, int64_t strideTau, T *__restrict__ B,
                                                                  int64_t ldb, int64_t strideB, int n, int nrhs, int64_t batch,
                                                                  int32_t *__restrict__ info_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int CB = 8, SUBS = 4, MAXR = 16;               // 8 columns x 4 row groups; 4 * 16 = 64 rows at most
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int jb = lane % CB, sub = lane / CB;
    const int P = n + 1;
    T *S = reinterpret_cast<T *>(smem_raw) + (size_t)warp * (P * n + 2 * n);
    T *taus = S + P * n, *rdiag = taus + n;
    const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
    for (int64_t mi = gw; mi < batch; mi += nw) {
        const T *Am = A + mi * strideA;
        __syncwarp();
        for (int idx = lane; idx < n * n; idx += 32) S[(idx / n) * P + (idx % n)] = Am[(int64_t)(idx / n) * lda + (idx % n)];
        for (int i = lane; i < n; i += 32) taus[i] = tau[mi * strideTau + i];
        __syncwarp();
        int bad = 0;
        if constexpr (MODE == 2) {
            for (int c = lane; c < n; c += 32) rdiag[c] = T(1) / S[c * P + c];
            for (int c = n - 1; c >= 0; --c)
                if (S[c * P + c] == T(0)) bad = c + 1;       // first zero on the diagonal, 1-based
            __syncwarp();
        }
        for (int c0 = 0; c0 < nrhs; c0 += CB) {
            const int col = c0 + jb;
            const bool cvalid = col < nrhs;
            T *Bc = B + mi * strideB + (int64_t)(cvalid ? col : 0) * ldb;
            T b[MAXR];
#pragma unroll
            for (int t = 0; t < MAXR; ++t) {
                const int i = sub + SUBS * t;
                b[t] = (cvalid && i < n) ? Bc[i] : T(0);
            }
            for (int kk = 0; kk < n; ++kk) {
                const int k = (MODE == 0) ? kk : n - 1 - kk;
                const T *vk = S + k * P;
                const T tk = taus[k];
                T part = T(0);
#pragma unroll
                for (int t = 0; t < MAXR; ++t) {
                    const int i = sub + SUBS * t;
                    if (i < k) part = t_fma(vk[i], b[t], part);
                    else if (i == k) part += b[t];
                }
                part += __shfl_xor_sync(FULL, part, CB);
                part += __shfl_xor_sync(FULL, part, 2 * CB);
                const T s = tk * part;
#pragma unroll
                for (int t = 0; t < MAXR; ++t) {
                    const int i = sub + SUBS * t;
                    if (i < k) b[t] = t_fma(-s, vk[i], b[t]);
                    else if (i == k) b[t] -= s;
                }
            }
            if constexpr (MODE == 2) {
                for (int c = 0; c < n; ++c) {
                    const int to = c / SUBS, so = c % SUBS;
                    T mine = T(0);
#pragma unroll
                    for (int t = 0; t < MAXR; ++t)
                        if (t == to) mine = b[t];
                    const T xc = __shfl_sync(FULL, mine, jb + CB * so) * rdiag[c];
                    const T *lc = S + c * P;
#pragma unroll
                    for (int t = 0; t < MAXR; ++t) {
                        const int i = sub + SUBS * t;
                        if (i == c) b[t] = xc;
                        else if (i > c && i < n) b[t] = t_fma(-lc[i], xc, b[t]);
                    }
                }
            }
            if (cvalid) {
#pragma unroll
                for (int t = 0; t < MAXR; ++t) {
                    const int i = sub + SUBS * t;
                    if (i < n) Bc[i] = b[t];
                }
            }
        }
        if (MODE == 2 && info_out != nullptr && lane == 0) info_out[mi] = bad;
    }
}

template <typename T, int MODE>
static int launch_apply_generic(qlb200_ctx *h, const T *A, int64_t lda, int64_t strideA, const T *tau, int64_t strideTau,
                                T *B, int64_t ldb, int64_t strideB, int n, int nrhs, int64_t batch, int32_t *info_out) {
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * ((size_t)(n + 1) * n + 2 * n) * sizeof(T);
    auto kern = apply_generic_kernel<T, MODE, WARPS>;
    int occ_unused = 0;
    int rcs = kernel_ready(h, kern, WARPS * 32, smem, &occ_unused);
    if (rcs) return rcs;
    int64_t blocks = (batch + WARPS - 1) / WARPS;
    const int64_t cap = (int64_t)h->sm_count * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, WARPS * 32, smem, h->stream>>>(A, lda, strideA, tau, strideTau, B, ldb, strideB, n, nrhs, batch, info_out);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

template <typename T, int N, int MODE>
static int dispatch_apply_nr(qlb200_ctx *h, const T *A, int64_t lda, int64_t strideA, const T *tau, int64_t strideTau,
                             T *B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch, int32_t *info_out, int nr = N) {
    // NR = columns of B handled per pass (a power of two <= 8; a configuration whose row blocks would not be whole vectors
    // falls through to the next wider one, which masks the missing columns)
    if constexpr (ApplyCfg<T, N, 1, MODE>::VALID) {
        if (nrhs <= 1) return launch_apply_batched<T, N, 1, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, nr);
    }
    if constexpr (ApplyCfg<T, N, 2, MODE>::VALID) {
        if (nrhs <= 2) return launch_apply_batched<T, N, 2, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, nr);
    }
    if constexpr (ApplyCfg<T, N, 4, MODE>::VALID) {
        if (nrhs <= 4) return launch_apply_batched<T, N, 4, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, nr);
    }
    static_assert(ApplyCfg<T, N, 8, MODE>::VALID, "the widest configuration must exist");
    return launch_apply_batched<T, N, 8, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, nr);
}

template <typename T, int MODE>
static int dispatch_apply_n(qlb200_ctx *h, int64_t n, const T *A, int64_t lda, int64_t strideA, const T *tau,
                            int64_t strideTau, T *B, int64_t ldb, int64_t strideB, int nrhs, int64_t batch,
                            int32_t *info_out) {
    switch (n) {
        case 8: return dispatch_apply_nr<T, 8, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
        case 16: return dispatch_apply_nr<T, 16, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
        case 32: return dispatch_apply_nr<T, 32, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
        case 64:
            if (h->opt_bvar != 9) return dispatch_apply_nr<T, 64, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out);
            break;
    }
    // other orders up to 32: the next larger tile, zero-padded in front (see the factorization)
    if (n >= 1 && n < 8) return dispatch_apply_nr<T, 8, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, (int)n);
    if (n > 8 && n < 16) return dispatch_apply_nr<T, 16, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, (int)n);
    if (n > 16 && n < 32) return dispatch_apply_nr<T, 32, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, (int)n);
    if (n > 32 && n < 64 && h->opt_bvar != 9) return dispatch_apply_nr<T, 64, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, nrhs, batch, info_out, (int)n);
    if (n >= 1 && n <= 64)
        return launch_apply_generic<T, MODE>(h, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)n, nrhs, batch, info_out);
    return qlb_fail(h, -2, "batched apply: n must be in [1,64]%s");
}

}  // namespace qlb

#include "batched_wy.cuh"

// ---------------------------------------------------------------------------------------------
// typed internal entry points used by capi.cu
// ---------------------------------------------------------------------------------------------
// (the Makefile compiles this file twice, -DQLB_BATCHED_ONLY_F64 / -DQLB_BATCHED_ONLY_F32: the two element types' kernel
//  instantiations build in parallel)
#ifndef QLB_BATCHED_ONLY_F32
int qlb_geqlf_batched_f64(qlb200_ctx *h, int64_t n, double *A, int64_t lda, int64_t strideA, double *tau,
                          int64_t strideTau, int64_t batch) {
    switch (n) {
        case 8: return qlb::launch_geqlf_batched<double, 8, 1, 4, 6>(h, A, lda, strideA, tau, strideTau, batch);
        case 16: return qlb::launch_geqlf_batched<double, 16, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch);
        case 32:
            if (h->opt_bvar == 8) return qlb::launch_geqlf_wy<double, 32>(h, A, lda, strideA, tau, strideTau, batch, 32);      // measurement only
            if (h->opt_bvar == 1) return qlb::launch_geqlf_batched<double, 32, 2, 4, 2>(h, A, lda, strideA, tau, strideTau, batch);
            // One kernel for small batches: below ~48k matrices (the shard of an 8-GPU run) the second kernel's fill/drain and
            // the two extra launches cost more than the split saves (32768 matrices: 209 vs 220 us; 262144: 1519 vs 1446 us).
            if (h->opt_bvar == 2 || (h->opt_bvar == 0 && batch < 49152))       // bvar 3 forces the split at any batch (tests)
                return qlb::launch_geqlf_batched<double, 32, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch);
            return qlb::launch_geqlf_batched_split32<double, 4, 3, 3>(h, A, lda, strideA, tau, strideTau, batch);
    }
    // every other order up to 32: the next larger register tile, the matrix in its bottom-right corner (zero padding, the
    // padded steps are not taken)
    if (n >= 1 && n < 8) return qlb::launch_geqlf_batched<double, 8, 1, 4, 6>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    if (n > 8 && n < 16) return qlb::launch_geqlf_batched<double, 16, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    if (n > 16 && n < 32) return qlb::launch_geqlf_batched<double, 32, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    // 32 < n <= 64: matrix in shared memory, panel in registers, DMMA trailing update (batched_wy.cuh); bvar 9 = literal ?geql2 kernel
    if (h->opt_bvar != 9) {
        if (n > 32 && n <= 40) return qlb::launch_geqlf_wy<double, 40>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
        if (n > 40 && n <= 48) return qlb::launch_geqlf_wy<double, 48>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
        if (n > 48 && n <= 56) return qlb::launch_geqlf_wy<double, 56>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
        if (n > 56 && n <= 64) return qlb::launch_geqlf_wy<double, 64>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    }
    if (n >= 1 && n <= 64) return qlb::launch_geql2_generic<double>(h, A, lda, strideA, tau, strideTau, (int)n, batch);
    return qlb_fail(h, -2, "batched QL: n must be in [1,64]%s");
}
#endif
#ifndef QLB_BATCHED_ONLY_F64
int qlb_geqlf_batched_f32(qlb200_ctx *h, int64_t n, float *A, int64_t lda, int64_t strideA, float *tau,
                          int64_t strideTau, int64_t batch) {
    switch (n) {
        case 8: return qlb::launch_geqlf_batched<float, 8, 1, 4, 8>(h, A, lda, strideA, tau, strideTau, batch);
        case 16:
            if (h->opt_bvar == 5) return qlb::launch_geqlf_batched<float, 16, 2, 4, 6>(h, A, lda, strideA, tau, strideTau, batch);      // experiment
            if (h->opt_bvar == 4) return qlb::launch_geqlf_batched<float, 16, 2, 4, 4>(h, A, lda, strideA, tau, strideTau, batch);      // experiment
            return qlb::launch_geqlf_batched<float, 16, 2, 4, 5>(h, A, lda, strideA, tau, strideTau, batch);
        case 32:
            if (h->opt_bvar == 2) return qlb::launch_geqlf_batched<float, 32, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch);
            return qlb::launch_geqlf_batched_split32<float, 4, 3, 5>(h, A, lda, strideA, tau, strideTau, batch);
    }
    if (n >= 1 && n < 8) return qlb::launch_geqlf_batched<float, 8, 1, 4, 8>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    if (n > 8 && n < 16) return qlb::launch_geqlf_batched<float, 16, 2, 4, 5>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    if (n > 16 && n < 32) return qlb::launch_geqlf_batched<float, 32, 2, 4, 3>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    // 32 < n <= 64: the Float64 shared-memory/DMMA kernel with the matrix widened while it is staged (bvar 9: literal ?geql2)
    if (h->opt_bvar != 9) {
        if (n > 32 && n <= 40) return qlb::launch_geqlf_wy<float, 40>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
        if (n > 40 && n <= 48) return qlb::launch_geqlf_wy<float, 48>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
        if (n > 48 && n <= 56) return qlb::launch_geqlf_wy<float, 56>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
        if (n > 56 && n <= 64) return qlb::launch_geqlf_wy<float, 64>(h, A, lda, strideA, tau, strideTau, batch, (int)n);
    }
    if (n >= 1 && n <= 64) return qlb::launch_geql2_generic<float>(h, A, lda, strideA, tau, strideTau, (int)n, batch);
    return qlb_fail(h, -2, "batched QL: n must be in [1,64]%s");
}
#endif
#ifndef QLB_BATCHED_ONLY_F32
// mode 0: Q*B, 1: Q'*B, 2: L^{-1} Q' B
int qlb_apply_batched_f64(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const double *A, int64_t lda,
                          int64_t strideA, const double *tau, int64_t strideTau, double *B, int64_t ldb,
                          int64_t strideB, int64_t batch, int32_t *info_out) {
    if (mode == 0) return qlb::dispatch_apply_n<double, 0>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    if (mode == 1) return qlb::dispatch_apply_n<double, 1>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    return qlb::dispatch_apply_n<double, 2>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
}
#endif
#ifndef QLB_BATCHED_ONLY_F64
int qlb_apply_batched_f32(qlb200_ctx *h, int mode, int64_t n, int64_t nrhs, const float *A, int64_t lda,
                          int64_t strideA, const float *tau, int64_t strideTau, float *B, int64_t ldb,
                          int64_t strideB, int64_t batch, int32_t *info_out) {
    if (mode == 0) return qlb::dispatch_apply_n<float, 0>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    if (mode == 1) return qlb::dispatch_apply_n<float, 1>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
    return qlb::dispatch_apply_n<float, 2>(h, n, A, lda, strideA, tau, strideTau, B, ldb, strideB, (int)nrhs, batch, info_out);
}
#endif
