// blocked.cu -- device-side drivers of the single-matrix hot path (all pointers are DEVICE pointers,
// everything is queued on h->stream and returns without synchronizing):
//
//   qlb_dgeqlf_dev    ql!(A): blocked Householder QL, packed output identical in layout to LAPACK
//                     ?geqlf / the reference's QL.factors + QL.tau (reference src/ql.jl:72, 37-45).
//                     Panels of nb columns from the right; each panel = strips of 16 columns factored
//                     by the cluster kernel (panel.cu) + in-panel compact-WY updates; then T from the
//                     Gram matrix (larft) and the trailing update C -= V (T' (V' C))' as three DMMA
//                     GEMMs (gemm.cu).  SURVEY.md A.2 spells out the backward/columnwise conventions.
//   qlb_dormql_dev    lmul!/rmul! with Q or Q' (reference src/ql.jl:321-435) as blocked ?ormql
//                     (SURVEY.md A.3): T rebuilt per block from V and tau, same GEMMs.
//   qlb_dqlsolve_dev  ldiv!(F,B), square (reference src/ql.jl:440-447,467): Q'B then a blocked
//                     forward substitution with L = tril(factors).
//   qlb_dgerqf_dev    rq!(A) (reference src/rq.jl:401) through gerqf(A) = geqlf(A')' (SURVEY.md A.4).
#include "common.cuh"
#include <vector>
#include "internal.h"

namespace qlb {

static inline int64_t even_up(int64_t x) { return (x + 1) & ~(int64_t)1; }

// Split-K factor for a GEMM with `tiles` output tiles and reduction length K: fill the SMs, keep
// chunks >= minchunk, never more than maxS partials.
static int choose_splits(int sm, int64_t tiles, int K, int minchunk, int maxS) {
    if (tiles <= 0) return 1;
    int smax = K / minchunk;
    if (smax > maxS) smax = maxS;
    if (smax < 1) smax = 1;
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= smax; ++s) {
        const double work = (double)tiles * s;
        const double waves = (double)((int64_t)((work + sm - 1) / sm));
        double eff = work / (waves * sm);
        eff *= 1.0 - 0.004 * (s - 1);          // partial-sum traffic penalty (relative: few-tile GEMMs must still split)
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    return best;
}

// Upper bound on the split-K factor of W = op(C) V for an operand with `other` columns (rows for side 'R'): thin
// operands (a few right-hand sides) give only a handful of tiles and need deep splits to occupy the machine.
static int max_w_splits(int64_t other) { return other >= 2048 ? 8 : (other >= 512 ? 32 : 128); }

using Workspace = ::QlbWorkspace;      // internal.h (shared with complex.cu)

// Carve the handle's workspace.  rows = reflector length bound, other = largest number of operand
// columns (rows for side 'R') a block reflector is applied to.  With `two` a second, independent set is
// carved for the look-ahead panel stream (its operands are at most nb wide).
static int get_workspace(qlb200_ctx *h, int64_t rows, int64_t other, int nb, int max_splits, Workspace *w,
                         Workspace *w2 = nullptr) {
    const int64_t ldv = even_up(rows), ldw = even_up(other), ldw2 = even_up(nb);
    const int64_t gram_part = (int64_t)160 * nb * nb;
    const int64_t strip_part = (int64_t)128 * even_up(nb) * 16;
    // split-K partials of W: the deepest split allowed for each operand width (max_w_splits)
    (void)max_splits;
    int64_t part = 0;
    {
        const int64_t widths[3] = {other, other < 2047 ? other : 2047, other < 511 ? other : 511};
        for (int i = 0; i < 3; ++i) {
            const int64_t cand = (int64_t)max_w_splits(widths[i]) * even_up(widths[i]) * nb;
            if (cand > part) part = cand;
        }
    }
    if (gram_part > part) part = gram_part;
    if (strip_part > part) part = strip_part;
    const int64_t part2 = gram_part > strip_part ? gram_part : strip_part;
    const int64_t fixed = ldv * nb + 2 * (int64_t)nb * nb + 256 + 64;
    int64_t total = fixed + 2 * ldw * nb + part;
    if (w2) total += fixed + 2 * ldw2 * nb + part2;
    int rc = qlb_reserve(h, &h->ws, &h->ws_bytes, (size_t)total * sizeof(double));
    if (rc) return rc;
    double *q = (double *)h->ws;
    auto carve = [&](Workspace *x, int64_t ldwx, int64_t partx) {
        x->Vw = q;  x->ldv = ldv;  q += ldv * nb;
        x->T = q;   x->ldt = nb;   q += (int64_t)nb * nb;
        x->G = q;                  q += (int64_t)nb * nb;
        x->Ts = q;                 q += 256 + 64;
        x->Wsum = q; x->ldw = ldwx; q += ldwx * nb;
        x->W2 = q;                 q += ldwx * nb;
        x->part = q; x->part_elems = partx; q += partx;
    };
    carve(w, ldw, part);
    if (w2) carve(w2, ldw2, part2);
    return 0;
}

// T (ib x ib) of the block reflector stored in ws.Vw[0:p, 0:ib] with scalars tau: Gram (split-K DMMA
// GEMM) -> reduce -> ?larft recurrence.
static int build_T(qlb200_ctx *h, cudaStream_t st, const Workspace &ws, int p, int ib, const double *tau) {
    int S = (p + 127) / 128;
    if (S > 148) S = 148;
    S = qlb_gemm_splits(p, S, 16);
    const int64_t stride = (int64_t)ib * ib;
    int rc = qlb_gemm(h, st, ib <= 16 ? 2 : -1, true, true, ib, ib, p, 1.0, ws.Vw, ws.ldv, ws.Vw, ws.ldv, 0.0,
                      S > 1 ? ws.part : ws.G, ib, S, stride);
    if (rc) return rc;
    if (S > 1) {
        rc = qlb_reduce_partials(h, st, ws.part, S, stride, ib, ib, ib, ws.G, ib);
        if (rc) return rc;
    }
    if (ib <= 128) return qlb_larft(h, st, ws.G, ib, tau, ib, ws.T, ws.ldt);
    // Two-level T for 128 < ib <= 256 (backward/columnwise): with the reflectors split into block 1 =
    // [0, h1) and block 2 = [h1, ib), H = H_2 H_1 = I - [V1 V2] [[T11, 0], [T21, T22]] [V1 V2]' with
    // T21 = -T22 (V2'V1) T11.  T11 and T22 by the ?larft recurrence on the diagonal Gram blocks, T21 by two small GEMMs
    // (the split-K partial buffer is free again after the reduction and holds the intermediate product).
    const int h1 = ib - 128, n2 = 128;
    rc = qlb_larft(h, st, ws.G, ib, tau, h1, ws.T, ws.ldt);
    if (rc) return rc;
    rc = qlb_larft(h, st, ws.G + h1 + (int64_t)h1 * ib, ib, tau + h1, n2, ws.T + h1 + (int64_t)h1 * ws.ldt, ws.ldt);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemset2DAsync(ws.T + (int64_t)h1 * ws.ldt, (size_t)ws.ldt * sizeof(double), 0, (size_t)h1 * sizeof(double), (size_t)n2, st));
    double *tmp = ws.part;                      // n2 x h1
    rc = qlb_gemm(h, st, -1, false, true, n2, h1, n2, 1.0, ws.T + h1 + (int64_t)h1 * ws.ldt, ws.ldt, ws.G + h1, ib, 0.0, tmp, n2, 1, 0);
    if (rc) return rc;
    return qlb_gemm(h, st, -1, false, true, n2, h1, h1, -1.0, tmp, n2, ws.T, ws.ldt, 0.0, ws.T + h1, ws.ldt, 1, 0);
}

// Apply the block reflector H = I - V T V' (V = ws.Vw[0:p,0:ib] or Vptr) to C.
//   left : C (p x nc, ldc)  <- (I - V Tm V') C      W = C' V (nc x ib)
//   right: C (mc x p, ldc)  <- C (I - V Tm V')      W = C V  (mc x ib)
// Tm = T (tt == false) or T' (tt == true) in W2 = W * Tm.
// wait_T: event after which T is valid (T may still be under construction on another stream while W is computed).
static int apply_block(qlb200_ctx *h, cudaStream_t st, const Workspace &ws, bool left, bool tt, const double *V,
                       int64_t ldv, const double *T, int64_t ldt, int p, int ib, double *C, int64_t ldc, int other,
                       cudaEvent_t wait_T = nullptr, int sms = 0, bool dense_T = false) {
    if (sms <= 0) sms = h->sm_count;               // SMs the stream can use (an SM partition in look-ahead mode)
    if (other <= 0 || p <= 0 || ib <= 0) return 0;
    // in-panel strip update: many thin partials, fused reduce*T (that kernel reads only the lower triangle of T: not for the
    // real block form of a complex T, whose 2x2 diagonal blocks reach above the diagonal)
    const bool small = ib <= 16 && other <= 256 && !dense_T;
    // "profile" option: the big block updates on the update partition (or the whole GPU) are the timed kernel; the panel
    // partition's GEMMs run concurrently with them and are not part of that roofline
    const bool timed = ib > 16 && 2 * sms >= h->sm_count;
    const int cfg_w = ib <= 16 ? 2 : -1;
    // 1. W partials: M = other, N = ib, K = p
    const int64_t tiles = (int64_t)((other + 127) / 128) * ((ib + 63) / 64);      // 128x64 tiles, 2 CTAs per SM
    int S;
    if (small) {
        S = (p + 127) / 128;
        if (S > 128) S = 128;
    } else if (h->opt_trail_splits > 0) {
        S = (int)h->opt_trail_splits;
    } else {
        S = choose_splits(2 * sms, tiles, p, other >= 2048 ? 512 : 128, max_w_splits(other));
    }
    S = qlb_gemm_splits(p, S, 16);
    const bool to_part = (S > 1 || small);
    const int64_t ldp = to_part ? even_up(other) : ws.ldw;      // partials are packed tightly
    const int64_t stride = ldp * (int64_t)ib;
    if (to_part && (int64_t)S * stride > ws.part_elems)
        return qlb_fail(h, 1000 + (int)cudaErrorMemoryAllocation, "partial buffer too small%s");
    double *Wdst = to_part ? ws.part : ws.Wsum;
    int rc;
    {
        QlbProfScope prof(h, st, timed ? 2.0 * other * ib * p : 0.0);
        rc = qlb_gemm(h, st, cfg_w, /*akc=*/left, /*bkc=*/true, other, ib, p, 1.0, C, ldc, V, ldv, 0.0, Wdst, ldp, S, stride);
    }
    if (rc) return rc;
    // 2. W2 = -(sum of partials) * Tm   (negated here, so that step 3 is a pure accumulation: alpha = beta = 1)
    if (wait_T) QLB_CUDA(h, cudaStreamWaitEvent(st, wait_T, 0));
    if (small && !tt) {
        rc = qlb_reduce_mult_t(h, st, ws.part, S, stride, other, ib, ldp, T, (int)ldt, ws.W2, ws.ldw);
        if (rc) return rc;
    } else {
        if (to_part) {
            rc = qlb_reduce_partials(h, st, ws.part, S, stride, other, ib, ldp, ws.Wsum, ws.ldw);
            if (rc) return rc;
        }
        rc = qlb_gemm(h, st, cfg_w, false, /*bkc=*/!tt, other, ib, ib, -1.0, ws.Wsum, ws.ldw, T, ldt, 0.0, ws.W2, ws.ldw, 1, 0);
        if (rc) return rc;
    }
    // 3. C += V W2' (left)  or  C += W2 V' (right), W2 already negated
    QlbProfScope prof(h, st, timed ? 2.0 * other * ib * p : 0.0);
    if (left) return qlb_gemm(h, st, -1, false, false, p, other, ib, 1.0, V, ldv, ws.W2, ws.ldw, 1.0, C, ldc, 1, 0);
    return qlb_gemm(h, st, -1, false, false, other, p, ib, 1.0, ws.W2, ws.ldw, V, ldv, 1.0, C, ldc, 1, 0);
}

// 32x32 tiled transpose: out (cols x rows, ldo) = in (rows x cols, ldi)'
__global__ void transpose_kernel(const double *__restrict__ in, int64_t ldi, int rows, int cols, double *__restrict__ out,
                                 int64_t ldo) {
    __shared__ double tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + threadIdx.x, c = c0 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[r + (int64_t)c * ldi];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + threadIdx.x, r = r0 + j;
        if (r < rows && c < cols) out[c + (int64_t)r * ldo] = tile[threadIdx.x][j];
    }
}

// Forward substitution with one tb x tb diagonal block of L (lower, non-unit): X = L^{-1} B in place.
// One CTA of TB threads; thread r owns row r of up to 8 right-hand sides at a time.
template <int TB>
__global__ void __launch_bounds__(TB) trsm_diag_kernel(const double *__restrict__ L, int64_t ldl, int tb, double *__restrict__ B,
                                                      int64_t ldb, int nrhs, int row_base, int32_t *__restrict__ info) {
    extern __shared__ __align__(16) double sm[];
    double *Ls = sm;                       // tb x tb, pitch TB+1
    __shared__ double xs[8];
    const int r = threadIdx.x;
    for (int idx = threadIdx.x; idx < tb * tb; idx += TB) {
        const int i = idx % tb, j = idx / tb;
        Ls[j * (TB + 1) + i] = (i >= j) ? L[i + (int64_t)j * ldl] : 0.0;
    }
    __syncthreads();
    if (info != nullptr && blockIdx.x == 0 && r < tb && Ls[r * (TB + 1) + r] == 0.0) atomicMin((int *)info, row_base + r + 1);
    for (int c0 = blockIdx.x * 8; c0 < nrhs; c0 += gridDim.x * 8) {       // CTAs share the right-hand sides
        const int nc = (nrhs - c0) < 8 ? (nrhs - c0) : 8;
        double b[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] = (r < tb && q < nc) ? B[r + (int64_t)(c0 + q) * ldb] : 0.0;
        for (int c = 0; c < tb; ++c) {
            if (r == c) {
                const double inv = 1.0 / Ls[c * (TB + 1) + c];
#pragma unroll
                for (int q = 0; q < 8; ++q) { b[q] *= inv; xs[q] = b[q]; }
            }
            __syncthreads();
            if (r > c && r < tb) {
                const double l = Ls[c * (TB + 1) + r];
#pragma unroll
                for (int q = 0; q < 8; ++q) b[q] = fma(-l, xs[q], b[q]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (r < tb && q < nc) B[r + (int64_t)(c0 + q) * ldb] = b[q];
    }
}

__global__ void set_int_kernel(int32_t *p, int32_t v) { *p = v; }

// 64-bit checksum of packed factors: sum (mod 2^64) over every entry of V (nq x k, any lda) and tau of a bijective mix of
// (bit pattern, position).  Any single modified entry changes the sum; the sum does not depend on lda, the addresses or
// the order of the additions.  It ties the T-factor cache to the CONTENT of the factors ql! produced.
__device__ __forceinline__ unsigned long long mix64(unsigned long long w, unsigned long long pos) {
    unsigned long long t = (w ^ (pos * 0x9E3779B97F4A7C15ull)) * 0xBF58476D1CE4E5B9ull;
    t ^= t >> 31;
    return t * 0x94D049BB133111EBull;
}
__global__ void set_u64_kernel(unsigned long long *p, unsigned long long v) { *p = v; }
__global__ void __launch_bounds__(256) checksum_kernel(const double *__restrict__ V, int64_t ldv, int nq, int k, const double *__restrict__ tau,
                                                       unsigned long long *__restrict__ out) {
    const int64_t total = (int64_t)nq * k + k;
    unsigned long long acc = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        double v;
        if (p < (int64_t)nq * k) {
            const int64_t j = p / nq, i = p - j * nq;
            v = V[i + j * ldv];
        } else {
            v = tau[p - (int64_t)nq * k];
        }
        acc += mix64((unsigned long long)__double_as_longlong(v), (unsigned long long)p);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc != 0) atomicAdd(out, acc);
}

}  // namespace qlb

using namespace qlb;

// The block-reflector machinery for complex.cu: a complex block reflector is applied as a REAL one of twice the size
// (V and T in their real 2x2-block form, C through its interleaved storage), see complex.cu.
int qlb_block_workspace(qlb200_ctx *h, int64_t rows, int64_t other, int nb, QlbWorkspace *w) {
    return get_workspace(h, rows, other, nb, max_w_splits(other), w);
}
int qlb_block_apply(qlb200_ctx *h, cudaStream_t st, const QlbWorkspace &ws, bool left, bool tt, const double *V, int64_t ldv,
                    const double *T, int64_t ldt, int p, int ib, double *C, int64_t ldc, int other) {
    return apply_block(h, st, ws, left, tt, V, ldv, T, ldt, p, ib, C, ldc, other, nullptr, 0, /*dense_T=*/true);
}
int qlb_block_gram(qlb200_ctx *h, cudaStream_t st, const QlbWorkspace &ws, int p, int ib) {
    // ws.G (ib x ib, ld ib) = V'V of ws.Vw[0:p, 0:ib]
    int S = (p + 127) / 128;
    if (S > 148) S = 148;
    S = qlb_gemm_splits(p, S, 16);
    const int64_t stride = (int64_t)ib * ib;
    int rc = qlb_gemm(h, st, ib <= 16 ? 2 : -1, true, true, ib, ib, p, 1.0, ws.Vw, ws.ldv, ws.Vw, ws.ldv, 0.0, S > 1 ? ws.part : ws.G, ib,
                      S, stride);
    if (rc) return rc;
    if (S > 1) rc = qlb_reduce_partials(h, st, ws.part, S, stride, ib, ib, ib, ws.G, ib);
    return rc;
}

static int factors_checksum(qlb200_ctx *h, cudaStream_t st, const double *V, int64_t ldv, int nq, int k, const double *tau,
                            unsigned long long *out) {
    set_u64_kernel<<<1, 1, 0, st>>>(out, 0x5851F42D4C957F2Dull ^ ((unsigned long long)nq << 32) ^ (unsigned long long)k);
    QLB_LAUNCH_CHECK(h);
    int64_t blocks = ((int64_t)nq * k + k + 2047) / 2048;
    const int64_t cap = (int64_t)h->sm_count * 8;
    if (blocks > cap) blocks = cap;
    checksum_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(V, ldv, nq, k, tau, out);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// Factor one panel (columns [c0, c1) of A, rows [0, p)) on stream st: strips from the right, each
// followed by its compact-WY update of the panel columns left of it; then the panel's T.
// Reflectors (clean copy) -> V (ldv), T -> Tp (ldt = nb).  `scr` provides Ts/G/W/partial scratch.
static int factor_panel(qlb200_ctx *h, cudaStream_t st, const Workspace &scr, double *A, int64_t lda, int m, int n, int k,
                        int c0, int c1, double *tau, double *V, int64_t ldv, double *Tp, int ldt, bool need_T, int sms = 0) {
    const int sb = 16, ib = c1 - c0, p = m - n + c1;
    int rc;
    int s1 = c1;
    while (s1 > c0) {
        const int s0 = (s1 - sb > c0) ? s1 - sb : c0;
        const int w = s1 - s0, ps = m - n + s1;
        // a single-strip panel writes its T straight into the panel's T
        double *Tdst = (ib <= sb) ? Tp : scr.Ts;
        const int ldtd = (ib <= sb) ? ldt : 16;
        rc = qlb_strip_factor(h, st, A + (int64_t)s0 * lda, lda, ps, w, tau + (s0 - (n - k)), V + (int64_t)(s0 - c0) * ldv, ldv, p,
                              Tdst, ldtd);
        if (rc) return rc;
        if (s0 > c0) {
            rc = apply_block(h, st, scr, true, false, V + (int64_t)(s0 - c0) * ldv, ldv, Tdst, ldtd, ps, w, A + (int64_t)c0 * lda,
                             lda, s0 - c0, nullptr, sms);
            if (rc) return rc;
        }
        s1 = s0;
    }
    if (need_T && ib > sb) {
        Workspace tmp = scr;
        tmp.Vw = V; tmp.ldv = ldv; tmp.T = Tp; tmp.ldt = ldt;
        rc = build_T(h, st, tmp, p, ib, tau + (c0 - (n - k)));
        if (rc) return rc;
    }
    return 0;
}

bool qlb_lookahead_active(const qlb200_ctx *h, int64_t m, int64_t n) {
    if (h->opt_lookahead == 1) return true;
    if (!(h->opt_lookahead == 2 && h->gc_ok)) return false;
    if ((m < n ? m : n) >= h->opt_la_min) return true;
    // wide matrices (the transposed tall-skinny RQ of configs[3]: 1024 x 65536): few, short panels, but every one of them is
    // followed by a long trailing update that hides the next panel (rq 65536 x 1024: 10.0 -> 8.1 ms)
    return m >= 1024 && n >= 8 * m && m * n >= (int64_t)4096 * 4096;
}

// Measured on B200 (tools/sweep_geqlf.py): with the look-ahead, 192-wide panels win at n = 16384 (241 vs 248 ms: the K = 192
// update GEMM runs at 27.4 instead of 25.9 TFLOP/s on the GEMM partition), tie at 12288 and lose below (the longer panels
// are exposed in the tail of the factorization); widths that are not multiples of 64 lose to tile quantisation.
int qlb_effective_nb(const qlb200_ctx *h, int64_t m, int64_t n) {
    if (h->opt_nb > 0) return (int)h->opt_nb;
    return (qlb_lookahead_active(h, m, n) && (m < n ? m : n) >= 14336) ? 192 : 128;
}

int qlb_dgeqlf_dev(qlb200_ctx *h, int64_t m64, int64_t n64, double *A, int64_t lda, double *tau) {
    const int m = (int)m64, n = (int)n64;
    const int k = m < n ? m : n;
    if (k == 0) return 0;
    const bool seq_host = (h->h2d.host && !h->h2d.la) || (h->d2h.host && !h->d2h.la);      // the chunk-joining host schedule without look-ahead
    const int nb = seq_host ? (h->opt_nb > 0 ? (int)h->opt_nb : 128) : qlb_effective_nb(h, m64, n64);
    cudaStream_t st = h->stream;
    const int npanels = (k + nb - 1) / nb;
    const bool la = npanels > 1 && !seq_host && qlb_lookahead_active(h, m64, n64);
    Workspace ws, ws2;
    int rc = get_workspace(h, m, n, nb, 0, &ws, &ws2);
    if (rc) return rc;
    auto panel_c1 = [&](int j) { return n - j * nb; };
    auto panel_c0 = [&](int j) { int c = n - (j + 1) * nb; return c > n - k ? c : n - k; };
    // T-factor cache: panel j's T goes straight into slot j (and is then also built for the last panel)
    h->tc_valid = false;
    double *tcache = nullptr;
    if (h->opt_tcache) {
        rc = qlb_reserve(h, &h->tcache, &h->tcache_bytes, (size_t)npanels * nb * nb * sizeof(double));
        if (rc) return rc;
        tcache = (double *)h->tcache;
    }
    auto finish_tcache = [&](double *tc) -> int {
        if (!tc) return 0;
        h->tc_key[0] = (int64_t)(uintptr_t)(A + (int64_t)(n - k) * lda);
        h->tc_key[1] = lda;
        h->tc_key[2] = (int64_t)(uintptr_t)tau;
        h->tc_key[3] = m;
        h->tc_key[4] = k;
        h->tc_key[5] = nb;
        h->tc_valid = true;
        if (h->opt_tcache == 1) {      // validated mode: the cache belongs to this CONTENT
            int r = factors_checksum(h, st, A + (int64_t)(n - k) * lda, lda, m, k, tau, h->tc_sum_dev);
            if (r) return r;
        }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone)
            QLB_CUDA(h, cudaEventRecord(h->ev_tc, st));      // (under capture the caller records it after the launch)
        return 0;
    };
    if (!la) {
        // The panel's T (Gram GEMM -> reduce -> ?larft, ~135 us of latency-bound work) is needed only for W2 = W T, not
        // for W = C'V: it is built on the high-priority side stream while the main stream already runs the big W GEMM.
        const bool overlap = h->opt_overlap_t != 0;
        // Host-pointer ql! (h->h2d.host set): the matrix arrives in chunks of CW columns, right to left, on the H2D stream.
        // A chunk JOINS the trailing update at a fixed panel index (deterministic schedule: join_step panels per chunk -- 1
        // measured best at n = 16384 --, and never later than the panel that needs its columns): the main stream waits for its arrival, applies the panels factored
        // so far to it (clean V extracted from the packed factors, T from the cache), and from then on it takes part in
        // the ordinary right-looking update.  Chunks that have not joined yet are simply not touched.
        const double *hsrc = (tcache != nullptr) ? h->h2d.host : nullptr;
        if (h->h2d.host && !hsrc) {       // no T cache: plain upload first
            QLB_CUDA(h, cudaMemcpy2DAsync(A, (size_t)lda * 8, h->h2d.host, (size_t)h->h2d.host_ld * 8, (size_t)m * 8, (size_t)n,
                                          cudaMemcpyHostToDevice, st));
        }
        int CW = (int)h->opt_chunk_cols;
        while ((n + CW - 1) / CW > 32) CW *= 2;
        const int nch = hsrc ? (n + CW - 1) / CW : 0;
        int issued = nch, next_join = nch - 1, joined_lo = hsrc ? n : 0;
        const int JSTEP = (int)h->opt_join_step;     // panels per joining chunk
        auto issue_chunk = [&](int c) -> int {
            const int lo = c * CW, hi = (lo + CW < n) ? lo + CW : n;
            if (c == nch - 1) QLB_CUDA(h, cudaEventRecord(h->ev_up[0], h->h2d_stream));
            QLB_CUDA(h, cudaMemcpy2DAsync(A + (int64_t)lo * lda, (size_t)lda * 8, hsrc + (int64_t)lo * h->h2d.host_ld,
                                          (size_t)h->h2d.host_ld * 8, (size_t)m * 8, (size_t)(hi - lo), cudaMemcpyHostToDevice,
                                          h->h2d_stream));
            QLB_CUDA(h, cudaEventRecord(h->ev_chunk[c], h->h2d_stream));
            if (c == nch - 1) {             // the first chunk on its way: its upload rate steers the host schedule choice
                QLB_CUDA(h, cudaEventRecord(h->ev_up[1], h->h2d_stream));
                h->up_bytes = (double)m * 8.0 * (double)(hi - lo);
            }
            return 0;
        };
        auto join_chunk = [&](int c, int nfact) -> int {      // nfact = panels factored so far
            const int lo = c * CW, hi = (lo + CW < n) ? lo + CW : n;
            QLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_chunk[c], 0));
            for (int i = 0; i < nfact; ++i) {
                const int c1i = panel_c1(i), c0i = panel_c0(i), ibi = c1i - c0i, pi = m - n + c1i;
                int r = qlb_extract_v(h, st, A + (int64_t)c0i * lda, lda, pi, ibi, pi - ibi, ws.Vw, ws.ldv);
                if (r) return r;
                r = apply_block(h, st, ws, true, false, ws.Vw, ws.ldv, tcache + (size_t)i * nb * nb, nb, pi, ibi, A + (int64_t)lo * lda, lda,
                                hi - lo);
                if (r) return r;
            }
            joined_lo = lo;
            return 0;
        };
        for (int j = 0; j < npanels; ++j) {
            const int c1 = panel_c1(j), c0 = panel_c0(j), ib = c1 - c0, p = m - n + c1;
            if (hsrc) {
                int cj = nch - 1 - j / JSTEP;                    // schedule: this chunk and everything right of it have joined
                if (c0 / CW < cj) cj = c0 / CW;                  // ... and in any case the chunk holding this panel's columns
                if (cj < 0) cj = 0;
                while (issued > 0 && issued - 1 >= cj - 2) {     // keep two chunks in flight ahead of the joins
                    --issued;
                    rc = issue_chunk(issued);
                    if (rc) return rc;
                }
                while (next_join >= cj) {
                    rc = join_chunk(next_join, j);
                    if (rc) return rc;
                    --next_join;
                }
            }
            double *Tj = tcache ? tcache + (size_t)j * nb * nb : ws.T;
            const bool big = ib > 16;                  // a single-strip panel gets its T from the strip kernel
            const bool need_T = c0 > 0 || tcache != nullptr;
            rc = factor_panel(h, st, ws, A, lda, m, n, k, c0, c1, tau, ws.Vw, ws.ldv, Tj, nb, need_T && big && !overlap);
            if (rc) return rc;
            if (h->d2h.host) {      // columns [c0, c1) are final now (host-pointer ql!): mark the moment ...
                // (the event slot was last used 8 panels ago; waits already queued keep the record they saw)
                QLB_CUDA(h, cudaEventRecord(h->ev_panel[j & 7], st));
            }
            cudaEvent_t wait_T = nullptr;
            if (need_T && big && overlap) {
                Workspace tmp = ws2;
                tmp.Vw = ws.Vw; tmp.ldv = ws.ldv; tmp.T = Tj; tmp.ldt = nb;
                if (c0 > 0) {
                    QLB_CUDA(h, cudaEventRecord(h->ev_a, st));
                    QLB_CUDA(h, cudaStreamWaitEvent(h->side_stream, h->ev_a, 0));
                    rc = build_T(h, h->side_stream, tmp, p, ib, tau + (c0 - (n - k)));
                    if (rc) return rc;
                    QLB_CUDA(h, cudaEventRecord(h->ev_b, h->side_stream));
                    wait_T = h->ev_b;
                } else {
                    rc = build_T(h, st, tmp, p, ib, tau + (c0 - (n - k)));
                    if (rc) return rc;
                }
            }
            if (c0 > joined_lo) {
                rc = apply_block(h, st, ws, true, false, ws.Vw, ws.ldv, Tj, nb, p, ib, A + (int64_t)joined_lo * lda, lda, c0 - joined_lo, wait_T);
                if (rc) return rc;
            } else if (wait_T) {
                QLB_CUDA(h, cudaStreamWaitEvent(st, wait_T, 0));      // nothing to update yet, but T must be complete before it is reused
            }
            if (h->d2h.host) {      // ... and send them home on the copy stream.  Queued AFTER the trailing update: a copy
                // to pageable memory blocks the host, and the GPU must have work to do meanwhile.
                QLB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_panel[j & 7], 0));
                QLB_CUDA(h, cudaMemcpy2DAsync(h->d2h.host + (int64_t)c0 * h->d2h.host_ld, (size_t)h->d2h.host_ld * 8,
                                              A + (int64_t)c0 * lda, (size_t)lda * 8, (size_t)(h->d2h.rows > 0 ? h->d2h.rows : m) * 8, (size_t)(c1 - c0),
                                              cudaMemcpyDeviceToHost, h->copy_stream));
            }
        }
        if (hsrc) {      // chunks left of every panel (wide matrices) or beyond the schedule: they join now
            while (issued > 0) {
                --issued;
                rc = issue_chunk(issued);
                if (rc) return rc;
            }
            while (next_join >= 0) {
                rc = join_chunk(next_join, npanels);
                if (rc) return rc;
                --next_join;
            }
        }
        if (h->d2h.host && n - k > 0) {      // wide matrix: the columns left of the reflectors are final only now
            QLB_CUDA(h, cudaEventRecord(h->ev_panel[0], st));
            QLB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_panel[0], 0));
            QLB_CUDA(h, cudaMemcpy2DAsync(h->d2h.host, (size_t)h->d2h.host_ld * 8, A, (size_t)lda * 8, (size_t)m * 8, (size_t)(n - k),
                                          cudaMemcpyDeviceToHost, h->copy_stream));
        }
        return finish_tcache(tcache);
    }
    // ---- look-ahead: panel j+1 is factored on the panel stream P while the update stream G applies panel j to the columns
    // left of panel j+1.  With the SM partitions (green contexts) P owns 16 SMs that G's GEMMs never occupy, so the cluster
    // strip kernels start at once; without them P is the high-priority side stream (correct, but the strips wait for SMs).
    // Vw alternates between the two workspaces; T goes straight into the T cache (one slot per panel) or alternates too.
    // G's GEMM scratch is ws, P's is ws2.
    const bool gc = h->gc_ok;
    cudaStream_t G = (gc && !h->opt_la_gemm_all) ? h->gc_gemm : st, P = gc ? h->gc_panel : h->side_stream;
    const int smsG = (gc && !h->opt_la_gemm_all) ? h->gc_gemm_sms : h->sm_count, smsP = gc ? h->gc_panel_sms : h->sm_count;
    cudaEvent_t evP = h->ev_a, evN = h->ev_b;      // "panel factored" (recorded on P), "next panel's columns updated" (on G)
    if (gc) {                                      // both partitions start after the caller's work on h->stream
        QLB_CUDA(h, cudaEventRecord(h->ev_gc[0], st));
        if (G != st) QLB_CUDA(h, cudaStreamWaitEvent(G, h->ev_gc[0], 0));
        QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_gc[0], 0));
    } else {
        QLB_CUDA(h, cudaEventRecord(h->ev_a, st));
        QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_a, 0));
    }
    double *Vb[2] = {ws.Vw, ws2.Vw}, *Tb[2] = {ws.T, ws2.T};
    auto Tslot = [&](int j) { return tcache ? tcache + (size_t)j * nb * nb : Tb[j & 1]; };
    // host destination given (split host-pointer ql!, capi.cu): the columns of a panel are final once it is factored -- they go
    // home on the copy stream behind an event of their own (ring of 8; the panel stream is never more than one panel ahead)
    auto send_home = [&](int j) -> int {
        if (!h->d2h.host) return 0;
        const int c0 = panel_c0(j), c1 = panel_c1(j);
        const int64_t rows = h->d2h.rows > 0 ? h->d2h.rows : m;
        QLB_CUDA(h, cudaEventRecord(h->ev_panel[j & 7], P));
        QLB_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_panel[j & 7], 0));
        QLB_CUDA(h, cudaMemcpy2DAsync(h->d2h.host + (int64_t)c0 * h->d2h.host_ld, (size_t)h->d2h.host_ld * 8, A + (int64_t)c0 * lda, (size_t)lda * 8,
                                      (size_t)rows * 8, (size_t)(c1 - c0), cudaMemcpyDeviceToHost, h->copy_stream));
        return 0;
    };
    // Host source (h->h2d.host with h->h2d.la; tall/square, T cache on): the matrix is uploaded in chunks of CW columns, right to
    // left, all queued on the H2D stream now.  Columns [joined_lo, n) take part in the factorization; a chunk JOINS on the
    // update stream at a fixed iteration (one chunk per panel, and never later than the panel that needs its columns -- a
    // deterministic schedule, so the result does not depend on transfer timing): G waits for its arrival, applies the panels
    // factored AND applied so far to it (clean V extracted from the packed factors, T from the cache), and from then on the
    // chunk is part of the ordinary trailing update.  Same arithmetic per column as the device-resident ql!.
    const double *hsrc = (h->h2d.host && h->h2d.la && tcache && m >= n) ? h->h2d.host : nullptr;
    int CW = (int)h->opt_chunk_cols;
    while ((n + CW - 1) / CW > 32) CW *= 2;
    const int nch = hsrc ? (n + CW - 1) / CW : 0;
    int joined_lo = hsrc ? n : 0, next_join = nch - 1;
    double *Vj = nullptr;
    if (hsrc) {
        rc = qlb_reserve(h, &h->joinV, &h->joinV_bytes, (size_t)ws.ldv * nb * sizeof(double));
        if (rc) return rc;
        Vj = (double *)h->joinV;
        QLB_CUDA(h, cudaStreamWaitEvent(h->h2d_stream, gc ? h->ev_gc[0] : h->ev_a, 0));      // after the caller's work on the device buffer
        for (int c = nch - 1; c >= 0; --c) {
            const int lo = c * CW, hi = (lo + CW < n) ? lo + CW : n;
            if (c == nch - 1) QLB_CUDA(h, cudaEventRecord(h->ev_up[0], h->h2d_stream));
            QLB_CUDA(h, cudaMemcpy2DAsync(A + (int64_t)lo * lda, (size_t)lda * 8, hsrc + (int64_t)lo * h->h2d.host_ld, (size_t)h->h2d.host_ld * 8,
                                          (size_t)m * 8, (size_t)(hi - lo), cudaMemcpyHostToDevice, h->h2d_stream));
            QLB_CUDA(h, cudaEventRecord(h->ev_chunk[c], h->h2d_stream));
            if (c == nch - 1) {
                QLB_CUDA(h, cudaEventRecord(h->ev_up[1], h->h2d_stream));
                h->up_bytes = (double)m * 8.0 * (double)(hi - lo);
            }
        }
    }
    // chunks down to column need_lo join; nfact = panels already applied to every joined column
    auto join_until = [&](int need_lo, int nfact) -> int {
        while (joined_lo > need_lo && next_join >= 0) {
            const int c = next_join--, lo = c * CW, hi = joined_lo;
            QLB_CUDA(h, cudaStreamWaitEvent(G, h->ev_chunk[c], 0));
            for (int i = 0; i < nfact; ++i) {
                const int c1i = panel_c1(i), c0i = panel_c0(i), ibi = c1i - c0i, pi = m - n + c1i;
                int r = qlb_extract_v(h, G, A + (int64_t)c0i * lda, lda, pi, ibi, pi - ibi, Vj, ws.ldv);
                if (r) return r;
                r = apply_block(h, G, ws, true, false, Vj, ws.ldv, Tslot(i), nb, pi, ibi, A + (int64_t)lo * lda, lda, hi - lo, nullptr, smsG);
                if (r) return r;
            }
            joined_lo = lo;
        }
        return 0;
    };
    if (hsrc) {      // the first panel's columns: P waits for their chunks (G's waits are queued by join_until)
        const int cfirst = panel_c0(0) / CW;
        rc = join_until(cfirst * CW, 0);
        if (rc) return rc;
        for (int c = nch - 1; c >= cfirst; --c) QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_chunk[c], 0));
    }
    rc = factor_panel(h, P, ws2, A, lda, m, n, k, panel_c0(0), panel_c1(0), tau, Vb[0], ws.ldv, Tslot(0), nb, panel_c0(0) > 0 || tcache != nullptr, smsP);
    if (rc) return rc;
    QLB_CUDA(h, cudaEventRecord(evP, P));
    rc = send_home(0);
    if (rc) return rc;
    const bool trace = getenv("QLB_LA_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s_) { if (trace) { cudaEvent_t e_; cudaEventCreate(&e_); cudaEventRecord(e_, s_); tev.push_back(e_); } };
    for (int j = 0; j < npanels; ++j) {
        const int c1 = panel_c1(j), c0 = panel_c0(j), ib = c1 - c0, p = m - n + c1;
        double *V = Vb[j & 1], *T = Tslot(j);
        if (hsrc) {      // panels 0 .. j-1 have been applied to [joined_lo, .): late chunks catch up with exactly those
            // schedule: one chunk per panel while the upload is the limit (the first JRAMP panels), then JSTEP per panel -- the
            // update stream is the limit from there on, and fewer, wider catch-up GEMMs run closer to the big ones' rate
            const int JRAMP = (int)h->opt_la_join_ramp, JS = (int)h->opt_la_join_chunks;
            const int nj = (j < JRAMP) ? j + 2 : JRAMP + 2 + (j - JRAMP) * JS + (JS - 1);
            int need = n - nj * CW;
            const int dnext = (j + 1 < npanels) ? panel_c0(j + 1) : 0;
            if (need > dnext) need = dnext;                                    // ... and always the next panel's columns
            if (j + 1 == npanels) need = 0;
            if (need < 0) need = 0;
            rc = join_until((need / CW) * CW, j);
            if (rc) return rc;
        }
        const int lo_all = joined_lo;               // columns [lo_all, c0) are updated by panel j
        // While the trailing update is long (many columns left), the panel partition is far ahead of the update partition:
        // it then also applies panel j to the columns of panel j+1 itself (16 SMs: slower, but off the update partition's
        // critical path -- that one runs nothing but the big GEMMs, back to back, with no G -> P -> G round trip per panel).
        // It must see panel j-1 applied to those columns first (event evR of the previous rest update).
        const bool pnarrow = gc && j + 1 < npanels && h->opt_la_pnarrow_min > 0 && panel_c0(j + 1) >= h->opt_la_pnarrow_min;
        if (hsrc && pnarrow) QLB_CUDA(h, cudaEventRecord(h->ev_join, G));      // (late chunks have caught up on G)
        mark(G);                                                // [5j+0] G free (previous rest update done)
        QLB_CUDA(h, cudaStreamWaitEvent(G, evP, 0));            // panel j (V, T) is complete
        if (j + 1 < npanels) {
            const int d0 = panel_c0(j + 1);          // next panel = columns [d0, c0)
            mark(G);                                            // [5j+1] G starts (panel j arrived)
            if (pnarrow) {
                if (j > 0) QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_rest[(j - 1) & 1], 0));
                if (hsrc) QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_join, 0));
                rc = apply_block(h, P, ws2, true, false, V, ws.ldv, T, nb, p, ib, A + (int64_t)d0 * lda, lda, c0 - d0, nullptr, smsP);
                if (rc) return rc;
            } else {
                rc = apply_block(h, G, ws, true, false, V, ws.ldv, T, nb, p, ib, A + (int64_t)d0 * lda, lda, c0 - d0, nullptr, smsG);
                if (rc) return rc;
                QLB_CUDA(h, cudaEventRecord(evN, G));
                QLB_CUDA(h, cudaStreamWaitEvent(P, evN, 0));
            }
            mark(G);                                            // [5j+2] next-panel update done
            mark(P);                                            // [5j+3] P starts panel j+1
            rc = factor_panel(h, P, ws2, A, lda, m, n, k, d0, c0, tau, Vb[(j + 1) & 1], ws.ldv, Tslot(j + 1), nb, d0 > 0 || tcache != nullptr, smsP);
            if (rc) return rc;
            mark(P);                                            // [5j+4] P done
            QLB_CUDA(h, cudaEventRecord(evP, P));
            rc = send_home(j + 1);
            if (rc) return rc;
            if (d0 > lo_all) {
                rc = apply_block(h, G, ws, true, false, V, ws.ldv, T, nb, p, ib, A + (int64_t)lo_all * lda, lda, d0 - lo_all, nullptr, smsG);
                if (rc) return rc;
            }
            QLB_CUDA(h, cudaEventRecord(h->ev_rest[j & 1], G));
        } else if (c0 > lo_all) {
            rc = apply_block(h, G, ws, true, false, V, ws.ldv, T, nb, p, ib, A + (int64_t)lo_all * lda, lda, c0 - lo_all, nullptr, smsG);
            if (rc) return rc;
        }
    }
    if (trace) {
        mark(G);
        cudaDeviceSynchronize();
        double gwait = 0, small = 0, ptime = 0, rest = 0;
        for (int j = 0; j + 1 < npanels; ++j) {
            float a_, b_, c_, d_;
            cudaEventElapsedTime(&a_, tev[5 * j], tev[5 * j + 1]);          // G waiting for panel j
            cudaEventElapsedTime(&b_, tev[5 * j + 1], tev[5 * j + 2]);      // next-panel update
            cudaEventElapsedTime(&c_, tev[5 * j + 3], tev[5 * j + 4]);      // panel j+1 on P
            cudaEventElapsedTime(&d_, tev[5 * j + 2], tev[5 * (j + 1)]);    // rest update
            gwait += a_; small += b_; ptime += c_; rest += d_;
            if (j % 8 == 0 || j + 2 >= npanels) fprintf(stderr, "it %3d  Gwait %.3f  next-panel upd %.3f  panel(P) %.3f  rest upd %.3f ms\n", j, a_, b_, c_, d_);
        }
        fprintf(stderr, "sum: Gwait %.2f  next-panel upd %.2f  panels %.2f  rest upd %.2f ms\n", gwait, small, ptime, rest);
        for (auto e_ : tev) cudaEventDestroy(e_);
    }
    // the caller's stream continues after both partitions (P's last event was already waited for by G)
    if (G != st) {
        QLB_CUDA(h, cudaEventRecord(h->ev_gc[1], G));
        QLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_gc[1], 0));
    }
    return finish_tcache(tcache);
}

// cache_mode: 0 = use the T cache if the pointers are the ones ql! just factored, 1 = use it (the caller validated the
// factors by fingerprint), -1 = never.
int qlb_dormql_dev(qlb200_ctx *h, char side, char trans, int64_t m64, int64_t n64, int64_t k64, const double *A,
                   int64_t lda, const double *tau, double *C, int64_t ldc, int cache_mode) {
    const int m = (int)m64, n = (int)n64, k = (int)k64;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    const int nq = left ? m : n, other = left ? n : m;
    if (m == 0 || n == 0 || k == 0) return 0;
    // blocks must line up with ql!'s panels for the T cache to be usable: follow the width it used when it may hit
    const int nb = h->opt_nb > 0 ? (int)h->opt_nb
                                 : ((h->opt_tcache && h->tc_valid && h->tc_key[3] == nq && h->tc_key[4] == k && cache_mode >= 0) ? (int)h->tc_key[5] : 128);
    cudaStream_t st = h->stream;
    Workspace ws;
    int rc = get_workspace(h, nq, other, nb, max_w_splits(other), &ws);
    if (rc) return rc;
    bool cached = h->opt_tcache && h->tc_valid && h->tc_key[3] == nq && h->tc_key[4] == k && h->tc_key[5] == nb && cache_mode >= 0;
    if (cached && h->opt_tcache == 2)            // trusted mode: the pointers ql! factored
        cached = h->tc_key[0] == (int64_t)(uintptr_t)A && h->tc_key[1] == lda && h->tc_key[2] == (int64_t)(uintptr_t)tau;
    if (cached && h->opt_tcache == 1 && cache_mode != 2) {          // validated mode: same content (checksum of every reflector entry and tau); cache_mode 2: the caller (the split host schedule) applies the factors ql! has just left
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
            (void)cudaGetLastError();
            cached = false;                      // no host round trip inside a capture: rebuild T
        } else {
            QLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_tc, 0));
            rc = factors_checksum(h, st, A, lda, nq, k, tau, h->tc_sum_dev + 1);
            if (rc) return rc;
            unsigned long long sums[2] = {0, 1};
            QLB_CUDA(h, cudaMemcpyAsync(sums, h->tc_sum_dev, sizeof(sums), cudaMemcpyDeviceToHost, st));
            QLB_CUDA(h, cudaStreamSynchronize(st));
            cached = sums[0] == sums[1];
        }
    }
    if (cached) QLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_tc, 0));
    const double *tcache = cached ? (const double *)h->tcache : nullptr;
    const int nblk = (k + nb - 1) / nb;
    // Blocks are cut from the RIGHT like ql!'s panels (block j = reflectors [k-(j+1)nb, k-j nb)), so that cached T factors
    // line up.  Q = H_k ... H_1: Q C and C Q' apply H_1 first (leftmost block first); Q' C and C Q the other way round.
    const bool ascending = (left && notran) || (!left && !notran);
    for (int bi = 0; bi < nblk; ++bi) {
        const int j = ascending ? nblk - 1 - bi : bi;
        const int i1 = k - j * nb;
        const int i = (i1 - nb > 0) ? i1 - nb : 0;
        const int ib = i1 - i;
        const int p = nq - k + i + ib;          // rows of V for this block
        rc = qlb_extract_v(h, st, A + (int64_t)i * lda, lda, p, ib, nq - k + i, ws.Vw, ws.ldv);
        if (rc) return rc;
        const double *T = ws.T;
        if (tcache) {
            T = tcache + (size_t)j * nb * nb;
        } else {
            rc = build_T(h, st, ws, p, ib, tau + i);
            if (rc) return rc;
        }
        // left: H C = C - V T V'C -> W2 = W T' (notrans), W T (trans); right: C H = C - (C V) T V' -> W T / W T'
        const bool tt = left ? notran : !notran;
        rc = apply_block(h, st, ws, left, tt, ws.Vw, ws.ldv, T, ws.ldt, p, ib, C, ldc, other);
        if (rc) return rc;
    }
    return 0;
}

// Back substitution with one tb x tb NON-unit upper-triangular diagonal block: X = U^{-1} B in place (the wide minimum-norm
// solve, qlb_dqlminnorm_dev).  dinfo: 1-based index (i0 + c + 1) of the first zero on the diagonal, kept by atomicMin.
template <int TB>
__global__ void __launch_bounds__(TB) trsm_upper_diag_kernel(const double *__restrict__ U, int64_t ldu, int tb, double *__restrict__ B,
                                                            int64_t ldb, int nrhs, int i0, int32_t *__restrict__ dinfo) {
    extern __shared__ __align__(16) double sm[];
    double *Us = sm;                       // tb x tb, pitch TB+1
    __shared__ double xs[8];
    __shared__ double rd[TB];
    const int r = threadIdx.x;
    for (int idx = threadIdx.x; idx < tb * tb; idx += TB) {
        const int i = idx % tb, j = idx / tb;
        Us[j * (TB + 1) + i] = (i < j) ? U[i + (int64_t)j * ldu] : 0.0;
    }
    if (r < tb) {
        const double d = U[r + (int64_t)r * ldu];
        rd[r] = 1.0 / d;
        if (d == 0.0 && dinfo && blockIdx.x == 0) atomicMin(dinfo, i0 + r + 1);
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
                for (int q = 0; q < 8; ++q) {
                    b[q] *= rd[c];
                    xs[q] = b[q];
                }
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

// W (m x n, ldw) = L of a wide QL (getL, reference src/ql.jl:214-217): tril(factors, n - m)
__global__ void extract_l_wide_kernel(const double *__restrict__ F, int64_t ldf, int m, int n, double *__restrict__ W, int64_t ldw) {
    const int64_t total = (int64_t)m * n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % m), j = (int)(idx / m);
        W[i + (int64_t)j * ldw] = (j <= i + (n - m)) ? F[i + (int64_t)j * ldf] : 0.0;
    }
}
// B (n x nrhs): rows [0, m) move to rows [n - m, n) (through the scratch S, m x nrhs), rows [0, n - m) become zero
__global__ void stash_rows_kernel(const double *__restrict__ B, int64_t ldb, int m, int nrhs, double *__restrict__ S) {
    const int64_t total = (int64_t)m * nrhs;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
        S[idx] = B[idx % m + (idx / m) * ldb];
}
__global__ void place_rows_kernel(double *__restrict__ B, int64_t ldb, int m, int n, int nrhs, const double *__restrict__ S) {
    const int64_t total = (int64_t)n * nrhs;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % n);
        const int64_t c = idx / n;
        B[i + c * ldb] = (i < n - m) ? 0.0 : S[(i - (n - m)) + c * m];
    }
}

// CTAs of a diagonal-block solve: one per 8 right-hand sides, at most one per SM (the 128 x 128 block fills shared memory)
int trsm_grid(qlb200_ctx *h, int nrhs) {
    int g = (nrhs + 7) / 8;
    if (g > h->sm_count) g = h->sm_count;
    return g < 1 ? 1 : g;
}

// Forward substitution with L = tril(A) (n x n, non-unit): B (n x nrhs) <- L^{-1} B, blocked by 128 rows
// (diagonal-block kernel + DMMA GEMM).  dinfo (may be null): receives the 1-based index of the first zero
// diagonal entry, or 0x7fffffff if there is none.
int qlb_trsm_lower_dev(qlb200_ctx *h, int n, int nrhs, const double *A, int64_t lda, double *B, int64_t ldb, int32_t *dinfo) {
    cudaStream_t st = h->stream;
    if (dinfo) {
        set_int_kernel<<<1, 1, 0, st>>>(dinfo, 0x7fffffff);
        QLB_LAUNCH_CHECK(h);
    }
    constexpr int TB = 128;
    static_assert(TB * (TB + 1) * 8 <= 227 * 1024, "diag block must fit in shared memory");
    const size_t smem = (size_t)TB * (TB + 1) * sizeof(double);
    if (!h->trsm_attr_done) {
        QLB_CUDA(h, cudaFuncSetAttribute(trsm_diag_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->trsm_attr_done = true;
    }
    for (int i = 0; i < n; i += TB) {                                            // src/ql.jl:467
        const int tb = (n - i) < TB ? (n - i) : TB;
        trsm_diag_kernel<TB><<<trsm_grid(h, nrhs), TB, smem, st>>>(A + i + (int64_t)i * lda, lda, tb, B + i, ldb, nrhs, i, dinfo);
        QLB_LAUNCH_CHECK(h);
        const int rem = n - i - tb;
        if (rem > 0) {
            int rc = qlb_gemm(h, st, nrhs <= 16 ? 2 : -1, false, true, rem, nrhs, tb, -1.0, A + (i + tb) + (int64_t)i * lda, lda,
                              B + i, ldb, 1.0, B + i + tb, ldb, 1, 0);
            if (rc) return rc;
        }
    }
    return 0;
}

int qlb_dqlsolve_dev(qlb200_ctx *h, int64_t n64, int64_t nrhs64, const double *A, int64_t lda, const double *tau, double *B,
                     int64_t ldb, int32_t *dinfo, int cache_mode) {
    const int n = (int)n64, nrhs = (int)nrhs64;
    if (n == 0 || nrhs == 0) return 0;
    int rc = qlb_dormql_dev(h, 'L', 'T', n, nrhs, n, A, lda, tau, B, ldb, cache_mode);      // src/ql.jl:445
    if (rc) return rc;
    return qlb_trsm_lower_dev(h, n, nrhs, A, lda, B, ldb, dinfo);
}

// SURVEY.md 8(f) N2, wide half: the minimum-norm solution of A x = b for m < n from the QL factors.  The reference's own branch
// (src/ql.jl:448-483: a QR-derived "trapezoid to triangular" loop) is @test_broken (test/test_ql.jl:148); this is the same idea
// done correctly with the kernels at hand: A = Q L, L = [L1 L2] (m x n, L2 lower triangular) -> c = Q'b; RQ of the trapezoid,
// L = [0 R] Z (qlb_dgerqf_dev on a copy of L); y = R^{-1} c; x = Z' [0; y].  B is n x nrhs: rows [0, m) hold b on entry, rows
// [0, n) hold x on exit (the buffer the reference's `\` allocates, src/ql.jl:496-501).  dinfo: first zero on R's diagonal.
int qlb_dqlminnorm_dev(qlb200_ctx *h, int64_t m64, int64_t n64, int64_t nrhs64, const double *A, int64_t lda, const double *tau, double *B,
                       int64_t ldb, int32_t *dinfo) {
    const int m = (int)m64, n = (int)n64, nrhs = (int)nrhs64;
    if (m == 0 || nrhs == 0) return 0;
    cudaStream_t st = h->stream;
    int rc = qlb_dormql_dev(h, 'L', 'T', m, nrhs, m, A + (int64_t)(n - m) * lda, lda, tau, B, ldb, -1);      // c = Q'b
    if (rc) return rc;
    const int64_t ldw = even_up(m);
    rc = qlb_reserve(h, &h->zws, &h->zws_bytes, ((size_t)ldw * n + (size_t)m + 2 + (size_t)m * nrhs) * sizeof(double));
    if (rc) return rc;
    double *W = (double *)h->zws, *tau2 = W + (size_t)ldw * n, *S = tau2 + m + 2;
    const unsigned gb = (unsigned)(h->sm_count * 8);
    extract_l_wide_kernel<<<gb, 256, 0, st>>>(A, lda, m, n, W, ldw);
    QLB_LAUNCH_CHECK(h);
    rc = qlb_dgerqf_dev(h, m, n, W, ldw, tau2);
    if (rc) return rc;
    if (dinfo) {
        set_int_kernel<<<1, 1, 0, st>>>(dinfo, 0x7fffffff);
        QLB_LAUNCH_CHECK(h);
    }
    constexpr int TB = 128;
    const size_t smem = (size_t)TB * (TB + 1) * sizeof(double);
    static_assert(TB * (TB + 1) * 8 + TB * 8 + 64 <= 227 * 1024, "diag block must fit in shared memory");
    QLB_CUDA(h, cudaFuncSetAttribute(trsm_upper_diag_kernel<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double *R = W + (int64_t)(n - m) * ldw;               // upper triangular m x m
    const int nblk = (m + TB - 1) / TB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int i = b * TB, tb = (m - i) < TB ? (m - i) : TB;
        trsm_upper_diag_kernel<TB><<<trsm_grid(h, nrhs), TB, smem, st>>>(R + i + (int64_t)i * ldw, ldw, tb, B + i, ldb, nrhs, i, dinfo);
        QLB_LAUNCH_CHECK(h);
        if (i > 0) {      // B[0:i] -= R[0:i, i:i+tb] * X[i:i+tb]
            rc = qlb_gemm(h, st, nrhs <= 16 ? 2 : -1, false, true, i, nrhs, tb, -1.0, R + (int64_t)i * ldw, ldw, B + i, ldb, 1.0, B, ldb, 1, 0);
            if (rc) return rc;
        }
    }
    stash_rows_kernel<<<gb, 256, 0, st>>>(B, ldb, m, nrhs, S);
    QLB_LAUNCH_CHECK(h);
    place_rows_kernel<<<gb, 256, 0, st>>>(B, ldb, m, n, nrhs, S);
    QLB_LAUNCH_CHECK(h);
    return qlb_dormrq_dev(h, 'L', 'T', n, nrhs, m, W, ldw, tau2, B, ldb);      // x = Z' [0; y]
}

// ormrq through index reversal: gerqf(A) = geqlf(A')' (SURVEY.md A.4), so the RQ reflectors (rows of the k x nq
// array A, LAPACK ?ormrq layout, reference src/rq.jl:130-137,183-202,234-242,272-291) are the QL reflectors of A'
// and Q_rq = Q_ql': op(Q_rq) = op'(Q_ql) with the transposition flipped.
int qlb_dormrq_dev(qlb200_ctx *h, char side, char trans, int64_t m64, int64_t n64, int64_t k64, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc) {
    const int m = (int)m64, n = (int)n64, k = (int)k64;
    if (m == 0 || n == 0 || k == 0) return 0;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    const int nq = left ? m : n;
    const int64_t ldt = even_up(nq);
    int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, (size_t)ldt * k * sizeof(double));
    if (rc) return rc;
    double *At = (double *)h->ws2;                  // nq x k: column i = reflector row i
    dim3 blk(32, 8);
    transpose_kernel<<<dim3((k + 31) / 32, (nq + 31) / 32), blk, 0, h->stream>>>(A, lda, k, nq, At, ldt);
    QLB_LAUNCH_CHECK(h);
    return qlb_dormql_dev(h, side, notran ? 'T' : 'N', m, n, k, At, ldt, tau, C, ldc, -1);
}

// ---------------------------------------------------------------------------------------------
// QR through index reversal (reference src/qr.jl:60-76 -> LAPACK.geqrf!, test/test_ql.jl:7-13):
// with J the exchange matrix, geqrf(A) = J_m geqlf(J_m A J_n) J_n with tau reversed, and Q_qr = J Q_ql J.
// ---------------------------------------------------------------------------------------------
// A <- J_m A J_n in place: element (i, j) trades places with (m-1-i, n-1-j); pairs are enumerated over the dense m x n index
__global__ void flip2d_inplace_kernel(double *__restrict__ A, int64_t lda, int m, int n) {
    const int64_t total = (int64_t)m * n, half = total / 2;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < half; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = p / m, i = p - j * m;
        double *x = A + i + j * lda, *y = A + (m - 1 - i) + (int64_t)(n - 1 - j) * lda;
        const double t = *x;
        *x = *y;
        *y = t;
    }
}
// out = J_m in J_n (out of place)
__global__ void flip2d_kernel(const double *__restrict__ in, int64_t ldi, int m, int n, double *__restrict__ out, int64_t ldo) {
    const int64_t total = (int64_t)m * n;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = p / m, i = p - j * m;
        out[i + j * ldo] = in[(m - 1 - i) + (int64_t)(n - 1 - j) * ldi];
    }
}
__global__ void reverse_inplace_kernel(double *__restrict__ x, int k) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k / 2; i += gridDim.x * blockDim.x) {
        const double t = x[i];
        x[i] = x[k - 1 - i];
        x[k - 1 - i] = t;
    }
}
__global__ void reverse_copy_kernel(const double *__restrict__ x, int k, double *__restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) y[i] = x[k - 1 - i];
}
// C <- J C (rows reversed, side 'L') or C <- C J (columns reversed, side 'R'), in place
__global__ void flip_rows_or_cols_kernel(double *__restrict__ C, int64_t ldc, int m, int n, int rows) {
    const int hm = rows ? m / 2 : m, hn = rows ? n : n / 2;
    const int64_t total = (int64_t)hm * hn;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = p / hm, i = p - j * hm;
        double *x = C + i + j * ldc, *y = rows ? C + (m - 1 - i) + j * ldc : C + i + (int64_t)(n - 1 - j) * ldc;
        const double t = *x;
        *x = *y;
        *y = t;
    }
}
static inline unsigned flip_grid(qlb200_ctx *h, int64_t work) {
    int64_t b = (work + 255) / 256;
    const int64_t cap = (int64_t)h->sm_count * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// B (m x n, ldb) <- rows rotated up by `shift`: new row i = old row (i + shift) mod m (through the second workspace)
__global__ void rotate_rows_kernel(const double *__restrict__ B, int64_t ldb, int m, int n, int shift, double *__restrict__ out,
                                   int64_t ldo) {
    const int64_t total = (int64_t)m * n;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = p / m, i = p - j * m;
        int64_t src = i + shift;
        if (src >= m) src -= m;
        out[i + j * ldo] = B[src + j * ldb];
    }
}
int qlb_rotate_rows_dev(qlb200_ctx *h, int64_t m, int64_t n, double *B, int64_t ldb, int64_t shift) {
    if (m == 0 || n == 0 || shift % m == 0) return 0;
    const int64_t ldo = even_up(m);
    int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, (size_t)ldo * n * sizeof(double));
    if (rc) return rc;
    double *out = (double *)h->ws2;
    rotate_rows_kernel<<<flip_grid(h, m * n), 256, 0, h->stream>>>(B, ldb, (int)m, (int)n, (int)(shift % m), out, ldo);
    QLB_LAUNCH_CHECK(h);
    QLB_CUDA(h, cudaMemcpy2DAsync(B, (size_t)ldb * 8, out, (size_t)ldo * 8, (size_t)m * 8, (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

// the two halves of geqrf around the QL factorization (capi.cu puts ql! -- graph replay included -- between them)
int qlb_qr_flip_in(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda) {
    if (m == 0 || n == 0) return 0;
    flip2d_inplace_kernel<<<flip_grid(h, m * n / 2), 256, 0, h->stream>>>(A, lda, (int)m, (int)n);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
int qlb_qr_flip_out(qlb200_ctx *h, int64_t m, int64_t n, double *A, int64_t lda, double *tau) {
    if (m == 0 || n == 0) return 0;
    const int k = (int)(m < n ? m : n);
    h->tc_valid = false;                         // the cached T factors belong to the flipped reflectors
    flip2d_inplace_kernel<<<flip_grid(h, m * n / 2), 256, 0, h->stream>>>(A, lda, (int)m, (int)n);
    QLB_LAUNCH_CHECK(h);
    reverse_inplace_kernel<<<flip_grid(h, k / 2), 256, 0, h->stream>>>(tau, k);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// op(Q_qr) C or C op(Q_qr) with the QR reflectors in the columns of the nq x k array A (LAPACK ?ormqr layout, what
// lmul!/rmul! with a QRPackedQ compute): Q_qr = J Q_ql J, so flip C, apply the QL reflectors J A J, flip C back.
int qlb_dormqr_dev(qlb200_ctx *h, char side, char trans, int64_t m64, int64_t n64, int64_t k64, const double *A, int64_t lda,
                   const double *tau, double *C, int64_t ldc) {
    const int m = (int)m64, n = (int)n64, k = (int)k64;
    if (m == 0 || n == 0 || k == 0) return 0;
    const bool left = (side == 'L' || side == 'l');
    const int nq = left ? m : n;
    cudaStream_t st = h->stream;
    const int64_t ldf = even_up(nq);
    int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, ((size_t)ldf * k + (size_t)k + 2) * sizeof(double));
    if (rc) return rc;
    double *F = (double *)h->ws2, *taur = F + (size_t)ldf * k;
    flip2d_kernel<<<flip_grid(h, (int64_t)nq * k), 256, 0, st>>>(A, lda, nq, k, F, ldf);
    QLB_LAUNCH_CHECK(h);
    reverse_copy_kernel<<<flip_grid(h, k), 256, 0, st>>>(tau, k, taur);
    QLB_LAUNCH_CHECK(h);
    flip_rows_or_cols_kernel<<<flip_grid(h, (int64_t)m * n / 2), 256, 0, st>>>(C, ldc, m, n, left ? 1 : 0);
    QLB_LAUNCH_CHECK(h);
    rc = qlb_dormql_dev(h, side, trans, m, n, k, F, ldf, taur, C, ldc, -1);
    if (rc) return rc;
    flip_rows_or_cols_kernel<<<flip_grid(h, (int64_t)m * n / 2), 256, 0, st>>>(C, ldc, m, n, left ? 1 : 0);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

// ldiv!(F::QR, B) for m >= n (the solve behind `qrunblocked(A) \\ b`, reference src/qr.jl:193-204): with F' = J F J the QL
// factors of J A J, A x = b  <=>  (J A J)(J x) = J b, so J x comes out of the QL solve (least squares when m > n) on the
// flipped right-hand side; after flipping back X = B[1:n, :] (LAPACK ?gels convention).
int qlb_dqrsolve_dev(qlb200_ctx *h, int64_t m64, int64_t n64, int64_t nrhs64, const double *A, int64_t lda, const double *tau,
                     double *B, int64_t ldb, int32_t *dinfo) {
    const int m = (int)m64, n = (int)n64, nrhs = (int)nrhs64;
    if (n == 0 || nrhs == 0) return 0;
    cudaStream_t st = h->stream;
    const int64_t ldf = even_up(m);
    int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, ((size_t)ldf * n + (size_t)n + 2) * sizeof(double));
    if (rc) return rc;
    double *F = (double *)h->ws2, *taur = F + (size_t)ldf * n;
    flip2d_kernel<<<flip_grid(h, (int64_t)m * n), 256, 0, st>>>(A, lda, m, n, F, ldf);
    QLB_LAUNCH_CHECK(h);
    reverse_copy_kernel<<<flip_grid(h, n), 256, 0, st>>>(tau, n, taur);
    QLB_LAUNCH_CHECK(h);
    flip_rows_or_cols_kernel<<<flip_grid(h, (int64_t)m * nrhs / 2), 256, 0, st>>>(B, ldb, m, nrhs, 1);
    QLB_LAUNCH_CHECK(h);
    rc = qlb_dormql_dev(h, 'L', 'T', m, nrhs, n, F, ldf, taur, B, ldb, -1);
    if (rc) return rc;
    rc = qlb_trsm_lower_dev(h, n, nrhs, F + (m - n), ldf, B + (m - n), ldb, dinfo);
    if (rc) return rc;
    flip_rows_or_cols_kernel<<<flip_grid(h, (int64_t)m * nrhs / 2), 256, 0, st>>>(B, ldb, m, nrhs, 1);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

int qlb_dgerqf_dev(qlb200_ctx *h, int64_t m64, int64_t n64, double *A, int64_t lda, double *tau) {
    const int m = (int)m64, n = (int)n64;
    if (m == 0 || n == 0) return 0;
    cudaStream_t st = h->stream;
    // A' (n x m) in a second buffer: gerqf(A) = geqlf(A')' with identical tau
    const int64_t ldt = even_up(n);
    int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, (size_t)ldt * m * sizeof(double));
    if (rc) return rc;
    double *At = (double *)h->ws2;
    dim3 blk(32, 8);
    transpose_kernel<<<dim3((m + 31) / 32, (n + 31) / 32), blk, 0, st>>>(A, lda, m, n, At, ldt);
    QLB_LAUNCH_CHECK(h);
    rc = qlb_dgeqlf_dev(h, n, m, At, ldt, tau);
    h->tc_valid = false;                         // those reflectors live in scratch memory
    if (rc) return rc;
    transpose_kernel<<<dim3((n + 31) / 32, (m + 31) / 32), blk, 0, st>>>(At, ldt, n, m, A, lda);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
