// gemm.cu -- K3: the FP64 tensor-core (DMMA, mma.sync.m16n8k8.f64) GEMM behind every Level-3 step of
// the blocked QL: the compact-WY trailing update ?larfb('L','T','B','C') that LAPACK ?geqlf runs
// inside the reference's ql! (reference src/ql.jl:72), the Gram matrix V'V for the
// T-factor (?larft), the blocked Q-apply (reference src/ql.jl:321-377 done as ?ormql)
// and the trsm updates of ldiv! (src/ql.jl:467).
//
//   D (M x N, column-major, ldd) = alpha * op(A) * op(B) + beta * D
//
// Operand layouts are template flags, so no operand is ever transposed in memory:
//   AKC: A(m,k) at A[k + m*lda] ("K-contiguous", BLAS transa='T')   else A[m + k*lda] (transa='N')
//   BKC: B(k,n) at B[k + n*ldb] ("K-contiguous", BLAS transb='N')   else B[n + k*ldb] (transb='T')
// Tiles are staged global -> shared with 16-byte cp.async (LDGSTS) through a STAGES-deep ring; the
// fragments are read with 16-byte LDS; tile pitches (== 4 mod 8 doubles for K-contiguous tiles, == 2
// mod 8 for M/N-contiguous ones) make the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups.  There is no FP64 tcgen05/wgmma kind on sm_100a: warp-level DMMA fed
// from registers is the FP64 tensor path of this architecture.
// Split-K (gridDim.z > 1) writes partial products to `D + z*split_stride`; partials are summed in
// a fixed order by reduce_partials_kernel, so results are bitwise reproducible.
#include "common.cuh"
#include "internal.h"

namespace qlb {

struct GemmParams {
    int M, N, K;
    const double *A;
    int64_t lda;
    const double *B;
    int64_t ldb;
    double *D;
    int64_t ldd;
    double alpha, beta;
    int kchunk;               // K range handled by one blockIdx.z (multiple of BK unless splits == 1)
    int64_t split_stride;     // elements between the partial outputs of consecutive z
    int pair_sync;            // > 0: the two N-tiles of an M-tile form a cluster and meet every pair_sync k-tiles (see launch_cfg)
};

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_, int MINB_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, WARPS = WARPS_M * WARPS_N, THREADS = WARPS * 32;
    static constexpr int MT = WM / 16, NT = WN / 8;
    static constexpr int PK = BK + 4;                     // pitch of a K-contiguous tile row (== 4 mod 8)
    static constexpr int PA_MC = BM + 2, PB_NC = BN + 2;  // pitch of an M/N-contiguous tile row (== 2 mod 8)
    static constexpr int A_ELEMS = (BM * PK > BK * PA_MC) ? BM * PK : BK * PA_MC;
    static constexpr int B_ELEMS = (BN * PK > BK * PB_NC) ? BN * PK : BK * PB_NC;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
};

// Stages one operand tile per k-step.  KC: tile is [ROWS][BK] (K contiguous), else [BK][ROWS] (row index
// contiguous).  VEC = 2: 16-byte cp.async (16-byte aligned base, even ld), VEC = 1: 8-byte.
// Every thread owns N fixed chunks of the tile; their shared-memory offsets and global pointers are
// computed once per tile and the pointers just advance by one k-step, so the steady-state cost is one
// cp.async + one 64-bit add per chunk (the predicated path is taken only by edge tiles / the K tail).
template <int ROWS, int BK, bool KC, int VEC, int THREADS>
struct TileLoader {
    static constexpr int PITCH = KC ? (BK + 4) : (ROWS + 2);
    static constexpr int INNER = KC ? BK : ROWS;         // contiguous extent of the tile
    static constexpr int OUTER = KC ? ROWS : BK;
    static constexpr int CPR = INNER / VEC;              // chunks per outer index
    static constexpr int CHUNKS = CPR * OUTER;
    static constexpr int N = (CHUNKS + THREADS - 1) / THREADS;
    static constexpr bool EXACT = (CHUNKS % THREADS) == 0;
    const double *gp[N];
    int so[N];
    int64_t kstep;
    int row0, k0;                                        // tile origin of the NEXT load
    bool rows_full;

    __device__ __forceinline__ void init(const double *g, int64_t ld, int row0_, int k0_, int rows_total, int tid) {
        row0 = row0_;
        k0 = k0_;
        rows_full = row0_ + ROWS <= rows_total;
        kstep = KC ? (int64_t)BK : (int64_t)BK * ld;
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const int idx = tid + q * THREADS;
            const int o = idx / CPR, i = (idx % CPR) * VEC;
            so[q] = o * PITCH + i;
            gp[q] = KC ? g + (int64_t)(row0_ + o) * ld + (k0_ + i) : g + (int64_t)(k0_ + o) * ld + (row0_ + i);
        }
    }
    __device__ __forceinline__ void load(double *s, const double *g, int64_t ld, int rows_total, int k_end, int tid) {
        if (rows_full && k0 + BK <= k_end) {
#pragma unroll
            for (int q = 0; q < N; ++q) {
                if (EXACT || tid + q * THREADS < CHUNKS) {
                    if constexpr (VEC == 2)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(s + so[q])), "l"(gp[q]) : "memory");
                    else
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(s + so[q])), "l"(gp[q]) : "memory");
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < N; ++q) {
                const int idx = tid + q * THREADS;
                if (!EXACT && idx >= CHUNKS) break;
                const int o = idx / CPR, i = (idx % CPR) * VEC;
                const int r = KC ? (row0 + o) : (row0 + i);      // operand row (m or n) index
                const int k = KC ? (k0 + i) : (k0 + o);
                const int inner_left = KC ? (k_end - k) : (rows_total - r);
                const bool outer_ok = KC ? (r < rows_total) : (k < k_end);
                int valid = outer_ok ? inner_left : 0;
                valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
                const double *src = valid > 0 ? gp[q] : g;
                if constexpr (VEC == 2)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s + so[q])), "l"(src), "r"(valid * 8)
                                 : "memory");
                else
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(s + so[q])), "l"(src), "r"(valid * 8)
                                 : "memory");
            }
        }
#pragma unroll
        for (int q = 0; q < N; ++q) gp[q] += kstep;
        k0 += BK;
    }
};

// ---------------------------------------------------------------------------------------------
// Fragment addressing.  The m16n8k8 DMMA fixes which (row, k) / (k, col) slot each lane feeds, but
// which MATRIX row / column / k index sits in a slot is free as long as A, B and the accumulators
// agree.  The maps below put the two values a lane needs next to each other in shared memory, so
// every fragment load is one 16-byte LDS and the accumulator pairs (c0,c2) / (c1,c3) are 16 contiguous
// bytes of a column of D:
//   MMA row g   -> tile row 2g,   MMA row g+8 -> tile row 2g+1         (within a 16-row m-tile)
//   MMA k   t   -> k 2t,          MMA k t+4   -> k 2t+1                (within a k8 group)
//   MMA col g of n-tile j -> tile column (j/2)*16 + 2g + (j&1)         (n-tiles paired)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lds128(const double *p, double &x, double &y) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(smem_u32(p)));
}

// a[i][0..3] for m-tiles i < MT of the warp tile at row wm0; As = stage tile, ks = k8 group
template <int MT, bool KC, int PITCH>
__device__ __forceinline__ void load_a_frags(double (&a)[MT][4], const double *As, int wm0, int ks, int g, int t) {
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int m = wm0 + i * 16 + 2 * g, k = ks * 8 + 2 * t;
        if constexpr (KC) {
            lds128(As + m * PITCH + k, a[i][0], a[i][2]);
            lds128(As + (m + 1) * PITCH + k, a[i][1], a[i][3]);
        } else {
            lds128(As + k * PITCH + m, a[i][0], a[i][1]);
            lds128(As + (k + 1) * PITCH + m, a[i][2], a[i][3]);
        }
    }
}
template <int NT, bool KC, int PITCH>
__device__ __forceinline__ void load_b_frags(double (&b)[NT][2], const double *Bs, int wn0, int ks, int g, int t) {
    static_assert(NT % 2 == 0, "n-tiles are paired");
#pragma unroll
    for (int jp = 0; jp < NT / 2; ++jp) {
        const int n = wn0 + jp * 16 + 2 * g, k = ks * 8 + 2 * t;
        if constexpr (KC) {
            lds128(Bs + n * PITCH + k, b[2 * jp][0], b[2 * jp][1]);
            lds128(Bs + (n + 1) * PITCH + k, b[2 * jp + 1][0], b[2 * jp + 1][1]);
        } else {
            lds128(Bs + k * PITCH + n, b[2 * jp][0], b[2 * jp + 1][0]);
            lds128(Bs + (k + 1) * PITCH + n, b[2 * jp][1], b[2 * jp + 1][1]);
        }
    }
}
// tile-relative row of accumulator registers (r = 0,1 -> row 2g; r = 2,3 -> row 2g+1) and column
__device__ __forceinline__ int acc_row(int wm0, int i, int g, int r) { return wm0 + i * 16 + 2 * g + (r >> 1); }
__device__ __forceinline__ int acc_col(int wn0, int j, int t, int r) { return wn0 + (j >> 1) * 16 + 2 * (2 * t + (r & 1)) + (j & 1); }

// acc = pre * D (tile origin m0,n0), D = alpha*acc: (c0,c2) and (c1,c3) are 16-byte pairs of a column.
template <int MT, int NT, bool UNIT>
__device__ __forceinline__ void acc_preload(double (&acc)[MT][NT][4], const double *D, int64_t ldd, int M, int N, int m0, int n0,
                                            int wm0, int wn0, int g, int t, double pre, bool vec_ok) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = m0 + acc_row(wm0, i, g, 0), n = n0 + acc_col(wn0, j, t, e);
                double x = 0.0, y = 0.0;
                if (n < N) {
                    const double *d = D + m + (int64_t)n * ldd;
                    if (vec_ok && m + 1 < M) {
                        const double2 v = *reinterpret_cast<const double2 *>(d);
                        x = v.x;
                        y = v.y;
                    } else {
                        if (m < M) x = d[0];
                        if (m + 1 < M) y = d[1];
                    }
                }
                acc[i][j][e] = UNIT ? x : pre * x;
                acc[i][j][2 + e] = UNIT ? y : pre * y;
            }
}
template <int MT, int NT, bool UNIT>
__device__ __forceinline__ void acc_store(const double (&acc)[MT][NT][4], double *D, int64_t ldd, int M, int N, int m0, int n0,
                                          int wm0, int wn0, int g, int t, double alpha, double beta_late, bool vec_ok) {
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = m0 + acc_row(wm0, i, g, 0), n = n0 + acc_col(wn0, j, t, e);
                if (n < N) {
                    double *d = D + m + (int64_t)n * ldd;
                    double x = UNIT ? acc[i][j][e] : alpha * acc[i][j][e], y = UNIT ? acc[i][j][2 + e] : alpha * acc[i][j][2 + e];
                    if (beta_late != 0.0) {          // alpha == 0: D = beta*D
                        if (m < M) x = fma(beta_late, d[0], x);
                        if (m + 1 < M) y = fma(beta_late, d[1], y);
                    }
                    if (vec_ok && m + 1 < M) {
                        *reinterpret_cast<double2 *>(d) = make_double2(x, y);
                    } else {
                        if (m < M) d[0] = x;
                        if (m + 1 < M) d[1] = y;
                    }
                }
            }
}

template <class Cfg, bool AKC, bool BKC, int VEC>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINB) gemm_kernel(const GemmParams p) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, MT = Cfg::MT, NT = Cfg::NT;
    constexpr int PA = AKC ? Cfg::PK : Cfg::PA_MC, PB = BKC ? Cfg::PK : Cfg::PB_NC;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % Cfg::WARPS_M) * Cfg::WM, wn0 = (warp / Cfg::WARPS_M) * Cfg::WN;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * p.kchunk;
    const int kend = (kbeg + p.kchunk < p.K) ? kbeg + p.kchunk : p.K;
    const int KT = (kend - kbeg + BK - 1) / BK;

    // Accumulators start from (beta/alpha)*D, loaded BEFORE the operand pipeline is primed, so the
    // latency of reading D overlaps the prologue instead of being exposed in the epilogue (the
    // rank-nb trailing update has only K = nb: every exposed round trip counts).
    double *D = p.D + (int64_t)blockIdx.z * p.split_stride;
    const bool preload = (p.beta != 0.0) && (p.alpha != 0.0);
    const double pre = preload ? p.beta / p.alpha : 0.0;
    double acc[MT][NT][4];
    const bool vec_ok = ((uintptr_t)D % 16 == 0) && (p.ldd % 2 == 0);
    if (preload) {
        acc_preload<MT, NT, false>(acc, D, p.ldd, p.M, p.N, m0, n0, wm0, wn0, g, t, pre, vec_ok);
    } else {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.0;
    }

    auto stage_a = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS; };
    auto stage_b = [&](int s) { return smem + (size_t)s * Cfg::STAGE_ELEMS + Cfg::A_ELEMS; };
    TileLoader<BM, BK, AKC, VEC, Cfg::THREADS> la;
    TileLoader<BN, BK, BKC, VEC, Cfg::THREADS> lb;
    la.init(p.A, p.lda, m0, kbeg, p.M, tid);
    lb.init(p.B, p.ldb, n0, kbeg, p.N, tid);
    auto issue = [&](int kt) {
        if (kt < KT) {
            const int s = kt % STAGES;
            la.load(stage_a(s), p.A, p.lda, p.M, kend, tid);
            lb.load(stage_b(s), p.B, p.ldb, p.N, kend, tid);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        if (p.pair_sync > 0 && kt % p.pair_sync == 0) {
            // both CTAs of the pair stream the SAME A tile: keeping them within pair_sync k-tiles of each other turns the
            // second read into an L2 hit (without it the pair drifts apart and DRAM traffic is 1.8x the algorithmic bytes)
            asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
        }
        issue(kt + STAGES - 1);
        const double *As = stage_a(kt % STAGES), *Bs = stage_b(kt % STAGES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            double a[MT][4], b[NT][2];
            load_a_frags<MT, AKC, PA>(a, As, wm0, ks, g, t);
            load_b_frags<NT, BKC, PB>(b, Bs, wn0, ks, g, t);
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma1688(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: D = alpha*acc (the beta*D term is already inside acc unless alpha == 0)
    acc_store<MT, NT, false>(acc, D, p.ldd, p.M, p.N, m0, n0, wm0, wn0, g, t, p.alpha, preload ? 0.0 : p.beta, vec_ok);
}

// out(M x N, ldo) = sum_{s < S} part[s*stride + m + n*ldp]   (fixed summation order)
__global__ void reduce_partials_kernel(const double *__restrict__ part, int S, int64_t stride, int M, int N, int64_t ldp,
                                       double *__restrict__ out, int64_t ldo) {
    const int64_t total = (int64_t)M * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx % M), n = (int)(idx / M);
        const double *src = part + m + (int64_t)n * ldp;
        double s = 0.0;
        for (int z = 0; z < S; ++z) s += src[(int64_t)z * stride];
        out[m + (int64_t)n * ldo] = s;
    }
}

// ---------------------------------------------------------------------------------------------
using CfgBig = GemmCfg<128, 128, 16, 64, 32, 4, 1>;     // 8 warps, 64x32 warp tiles, 1 CTA/SM
using CfgMid = GemmCfg<128, 64, 16, 32, 32, 3, 2>;      // 8 warps, 32x32 warp tiles, 2 CTAs/SM
using CfgSkinny = GemmCfg<128, 16, 16, 16, 16, 4, 2>;   // N <= 16 (strip updates): 8 warps of 16x16
using CfgBig32 = GemmCfg<128, 128, 32, 64, 32, 3, 1>;   // BK = 32 variant

template <class Cfg, bool AKC, bool BKC, int VEC>
static int launch_cfg(qlb200_ctx *h, int slot, const GemmParams &p, int splits, cudaStream_t st) {
    auto kern = gemm_kernel<Cfg, AKC, BKC, VEC>;
    if (!h->gemm_attr_done[slot]) {          // once per handle (= per device) and instantiation
        QLB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        h->gemm_attr_done[slot] = true;
    }
    dim3 grid((p.M + Cfg::BM - 1) / Cfg::BM, (p.N + Cfg::BN - 1) / Cfg::BN, splits);
    const int kt_total = (p.kchunk + Cfg::BK - 1) / Cfg::BK;
    if (h->opt_pair_sync > 0 && (grid.y == 2 || grid.y == 4) && kt_total >= 4 * h->opt_pair_sync) {
        GemmParams q = p;
        q.pair_sync = (int)h->opt_pair_sync;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
        cfg.dynamicSmemBytes = Cfg::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = grid.y;      // all N-tiles of an M-tile (2 for nb = 128, 4 for nb = 256) stream the same A tile
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        QLB_CUDA(h, cudaLaunchKernelEx(&cfg, kern, q));
        return 0;
    }
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

template <class Cfg>
static int launch_layout(qlb200_ctx *h, int cfg, bool akc, bool bkc, bool vec2, const GemmParams &p, int splits,
                         cudaStream_t st) {
#define QLB_GEMM_CASE(A_, B_)                                                                          \
    if (akc == A_ && bkc == B_) {                                                                      \
        const int slot = cfg * 8 + (A_ ? 4 : 0) + (B_ ? 2 : 0);                                        \
        return vec2 ? launch_cfg<Cfg, A_, B_, 2>(h, slot + 1, p, splits, st)                           \
                    : launch_cfg<Cfg, A_, B_, 1>(h, slot, p, splits, st);                              \
    }
    QLB_GEMM_CASE(true, true)
    QLB_GEMM_CASE(false, true)
    QLB_GEMM_CASE(false, false)
    QLB_GEMM_CASE(true, false)
#undef QLB_GEMM_CASE
    return -1;
}

}  // namespace qlb

// D = alpha*op(A)*op(B) + beta*D on stream `st`.  cfg: 0 big, 1 mid, 2 skinny, 3 big/BK32, -1 auto.
// splits > 1: partial products to D + z*split_stride (alpha = 1, beta = 0 enforced by the caller).
int qlb_gemm(qlb200_ctx *h, cudaStream_t st, int cfg, bool akc, bool bkc, int M, int N, int K, double alpha,
             const double *A, int64_t lda, const double *B, int64_t ldb, double beta, double *D, int64_t ldd, int splits,
             int64_t split_stride) {
    using namespace qlb;
    if (M <= 0 || N <= 0) return 0;
    if (cfg < 0) {
        if (N <= 16) cfg = 2;
        else if (h->opt_gemm_cfg >= 0) cfg = (int)h->opt_gemm_cfg;
        else cfg = 1;      // 128x64 tiles, 2 CTAs/SM: measured best for the skinny-K update and for long K
    }
    int bk = (cfg == 3) ? 32 : 16;
    if (splits < 1) splits = 1;
    int kchunk = K;
    if (splits > 1) {
        kchunk = ((K + splits - 1) / splits + bk - 1) / bk * bk;
        if (kchunk < bk) kchunk = bk;
        splits = (K + kchunk - 1) / kchunk;
        if (splits < 1) splits = 1;
    }
    GemmParams p{M, N, K, A, lda, B, ldb, D, ldd, alpha, beta, kchunk, split_stride};
    const bool vec2 = (((uintptr_t)A | (uintptr_t)B) % 16 == 0) && (lda % 2 == 0) && (ldb % 2 == 0);
    switch (cfg) {
        case 0: return launch_layout<CfgBig>(h, 0, akc, bkc, vec2, p, splits, st);
        case 1: return launch_layout<CfgMid>(h, 1, akc, bkc, vec2, p, splits, st);
        case 2: return launch_layout<CfgSkinny>(h, 2, akc, bkc, vec2, p, splits, st);
        case 3: return launch_layout<CfgBig32>(h, 3, akc, bkc, vec2, p, splits, st);
    }
    return qlb_fail(h, -2, "unknown gemm config%s");
}

// number of K-splits the split-K launch above will really use (callers size the partial buffer with it)
int qlb_gemm_splits(int K, int splits, int bk) {
    if (splits <= 1) return 1;
    int kchunk = ((K + splits - 1) / splits + bk - 1) / bk * bk;
    if (kchunk < bk) kchunk = bk;
    int s = (K + kchunk - 1) / kchunk;
    return s < 1 ? 1 : s;
}

int qlb_reduce_partials(qlb200_ctx *h, cudaStream_t st, const double *part, int S, int64_t stride, int M, int N,
                        int64_t ldp, double *out, int64_t ldo) {
    if (M <= 0 || N <= 0) return 0;
    const int64_t total = (int64_t)M * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > h->sm_count * 8) blocks = h->sm_count * 8;
    qlb::reduce_partials_kernel<<<blocks, 256, 0, st>>>(part, S, stride, M, N, ldp, out, ldo);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
