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
// pitch of every shared tile is == 4 (mod 16) doubles so the 8-byte fragment loads of a half-warp
// hit 16 distinct bank pairs.  There is no FP64 tcgen05/wgmma kind on sm_100a: warp-level DMMA fed
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
};

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_, int MINB_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN, WARPS = WARPS_M * WARPS_N, THREADS = WARPS * 32;
    static constexpr int MT = WM / 16, NT = WN / 8;
    static constexpr int PK = BK + 4;                     // pitch of a K-contiguous tile row
    static constexpr int PA_MC = BM + 4, PB_NC = BN + 4;  // pitch of an M/N-contiguous tile row
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
    static constexpr int PITCH = KC ? (BK + 4) : (ROWS + 4);
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
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = m0 + wm0 + i * 16 + g + (r >> 1) * 8;
                const int n = n0 + wn0 + j * 8 + 2 * t + (r & 1);
                acc[i][j][r] = (preload && m < p.M && n < p.N) ? pre * D[m + (int64_t)n * p.ldd] : 0.0;
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
        issue(kt + STAGES - 1);
        const double *As = stage_a(kt % STAGES), *Bs = stage_b(kt % STAGES);
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
            double a[MT][4], b[NT][2];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int m = wm0 + i * 16 + g, k = ks * 8 + t;
                if constexpr (AKC) {
                    a[i][0] = As[m * PA + k];
                    a[i][1] = As[(m + 8) * PA + k];
                    a[i][2] = As[m * PA + k + 4];
                    a[i][3] = As[(m + 8) * PA + k + 4];
                } else {
                    a[i][0] = As[k * PA + m];
                    a[i][1] = As[k * PA + m + 8];
                    a[i][2] = As[(k + 4) * PA + m];
                    a[i][3] = As[(k + 4) * PA + m + 8];
                }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int n = wn0 + j * 8 + g, k = ks * 8 + t;
                if constexpr (BKC) {
                    b[j][0] = Bs[n * PB + k];
                    b[j][1] = Bs[n * PB + k + 4];
                } else {
                    b[j][0] = Bs[k * PB + n];
                    b[j][1] = Bs[(k + 4) * PB + n];
                }
            }
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) dmma1688(acc[i][j], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: D = alpha*acc (the beta*D term is already inside acc)
    const double alpha = p.alpha;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = m0 + wm0 + i * 16 + g + (r >> 1) * 8;
                const int n = n0 + wn0 + j * 8 + 2 * t + (r & 1);
                if (m < p.M && n < p.N) {
                    double *d = D + m + (int64_t)n * p.ldd;
                    double v = alpha * acc[i][j][r];
                    if (!preload && p.beta != 0.0) v = fma(p.beta, *d, v);      // alpha == 0
                    *d = v;
                }
            }
}

// ---------------------------------------------------------------------------------------------
// Persistent rank-K update  D = alpha * A * B' + beta * D  for SHORT K (the trailing update of the
// blocked QL: K = nb).  A is M x K (M-contiguous), B is N x K (N-contiguous).  With only K/16 k-steps
// per tile the pipeline fill and the read of D dominate a tile-per-CTA kernel, so here two CTAs per SM
// each walk a list of tiles and (1) the cp.async operand ring runs continuously across tile
// boundaries, (2) the D tile of the NEXT tile is pulled into L2 (prefetch.global.L2) while the current
// tile is multiplied, so the accumulator preload at the next tile start is an L2 hit.
// ---------------------------------------------------------------------------------------------
struct RankKCfg {
    static constexpr int BM = 128, BN = 64, BK = 16, WM = 32, WN = 32, STAGES = 3, THREADS = 256;
    static constexpr int WARPS_M = BM / WM, MT = WM / 16, NT = WN / 8;
    static constexpr int PA = BM + 4, PB = BN + 4;          // pitches (doubles)
    static constexpr int A_ELEMS = BK * PA, B_ELEMS = BK * PB, STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
};

__global__ void __launch_bounds__(RankKCfg::THREADS, 2) rankk_persistent_kernel(const GemmParams p, const int tiles_m,
                                                                               const int ntiles) {
    using Cfg = RankKCfg;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, MT = Cfg::MT, NT = Cfg::NT;
    constexpr int PA = Cfg::PA, PB = Cfg::PB;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % Cfg::WARPS_M) * Cfg::WM, wn0 = (warp / Cfg::WARPS_M) * Cfg::WN;
    const int KT = (p.K + BK - 1) / BK;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_steps = my_tiles * KT;
    const double pre = p.beta / p.alpha;

    auto tile_origin = [&](int i, int &m0, int &n0) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        m0 = (tile % tiles_m) * BM;
        n0 = (tile / tiles_m) * BN;
    };
    TileLoader<BM, BK, false, 2, Cfg::THREADS> la;
    TileLoader<BN, BK, false, 2, Cfg::THREADS> lb;
    int iss_tile = 0, iss_kt = 0;          // issue-side position: runs STAGES-1 k-steps ahead, across tiles
    auto issue = [&](int step) {
        if (step < total_steps) {
            if (iss_kt == 0) {
                int m0, n0;
                tile_origin(iss_tile, m0, n0);
                la.init(p.A, p.lda, m0, 0, p.M, tid);
                lb.init(p.B, p.ldb, n0, 0, p.N, tid);
            }
            const int s = step % STAGES;
            double *As = smem + (size_t)s * Cfg::STAGE_ELEMS, *Bs = As + Cfg::A_ELEMS;
            la.load(As, p.A, p.lda, p.M, p.K, tid);
            lb.load(Bs, p.B, p.ldb, p.N, p.K, tid);
            if (++iss_kt == KT) {
                iss_kt = 0;
                ++iss_tile;
            }
        }
        cp_async_commit();
    };
    // pull the D tile of tile i into L2: 64 columns x 1 KiB = 512 lines of 128 bytes, two per thread
    auto prefetch_d = [&](int i) {
        if (i < my_tiles) {
            int m0, n0;
            tile_origin(i, m0, n0);
#pragma unroll
            for (int q = 0; q < (BN * BM / 16) / Cfg::THREADS; ++q) {
                const int line = tid + q * Cfg::THREADS;
                const int c = line / (BM / 16), r = (line % (BM / 16)) * 16;
                if (m0 + r < p.M && n0 + c < p.N) {
                    const double *ptr = p.D + m0 + r + (int64_t)(n0 + c) * p.ldd;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
                }
            }
        }
    };
    prefetch_d(0);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    int step = 0;
    for (int i = 0; i < my_tiles; ++i) {
        int m0, n0;
        tile_origin(i, m0, n0);
        prefetch_d(i + 1);
        double acc[MT][NT][4];
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = m0 + wm0 + a * 16 + g + (r >> 1) * 8, n = n0 + wn0 + b * 8 + 2 * t + (r & 1);
                    acc[a][b][r] = (m < p.M && n < p.N) ? pre * p.D[m + (int64_t)n * p.ldd] : 0.0;
                }
        for (int kt = 0; kt < KT; ++kt, ++step) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            issue(step + STAGES - 1);
            const double *As = smem + (size_t)(step % STAGES) * Cfg::STAGE_ELEMS, *Bs = As + Cfg::A_ELEMS;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                double a[MT][4], b[NT][2];
#pragma unroll
                for (int x = 0; x < MT; ++x) {
                    const int m = wm0 + x * 16 + g, k = ks * 8 + t;
                    a[x][0] = As[k * PA + m];
                    a[x][1] = As[k * PA + m + 8];
                    a[x][2] = As[(k + 4) * PA + m];
                    a[x][3] = As[(k + 4) * PA + m + 8];
                }
#pragma unroll
                for (int y = 0; y < NT; ++y) {
                    const int n = wn0 + y * 8 + g, k = ks * 8 + t;
                    b[y][0] = Bs[k * PB + n];
                    b[y][1] = Bs[(k + 4) * PB + n];
                }
#pragma unroll
                for (int x = 0; x < MT; ++x)
#pragma unroll
                    for (int y = 0; y < NT; ++y) dmma1688(acc[x][y], a[x], b[y]);
            }
        }
        // epilogue: plain stores (D's beta term is already in the accumulators)
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int m = m0 + wm0 + a * 16 + g + (r >> 1) * 8, n = n0 + wn0 + b * 8 + 2 * t + (r & 1);
                    if (m < p.M && n < p.N) p.D[m + (int64_t)n * p.ldd] = p.alpha * acc[a][b][r];
                }
    }
    cp_async_wait<0>();
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
        else cfg = (K <= 256 && N > 64) ? 1 : 0;      // short-K updates: 2 CTAs/SM hide the tile prologue
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
    // short-K rank update with enough tiles to keep every SM busy: persistent pipelined kernel
    if (!akc && !bkc && splits == 1 && K <= 512 && alpha != 0.0 && beta != 0.0 && vec2 && h->opt_rankk && h->opt_gemm_cfg < 0) {
        const int tiles_m = (M + RankKCfg::BM - 1) / RankKCfg::BM, tiles_n = (N + RankKCfg::BN - 1) / RankKCfg::BN;
        const int64_t ntiles = (int64_t)tiles_m * tiles_n;
        if (ntiles >= 4 * (int64_t)h->sm_count && ntiles < (1 << 30)) {
            if (!h->gemm_attr_done[63]) {
                QLB_CUDA(h, cudaFuncSetAttribute(rankk_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)RankKCfg::SMEM));
                h->gemm_attr_done[63] = true;
            }
            rankk_persistent_kernel<<<2 * h->sm_count, RankKCfg::THREADS, RankKCfg::SMEM, st>>>(p, tiles_m, (int)ntiles);
            QLB_LAUNCH_CHECK(h);
            return 0;
        }
    }
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
