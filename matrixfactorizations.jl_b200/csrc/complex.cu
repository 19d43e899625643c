// complex.cu -- SURVEY.md 8(f) N1: the QL hot path for ComplexF64 (interleaved re/im, column-major = Julia's
// Matrix{ComplexF64}): ql! (LAPACK zgeqlf conventions, what reference src/ql.jl:72 runs for Complex BlasFloat), lmul!/rmul!
// with Q or Q' (src/ql.jl:321-435, the conj placements at :336,366,368,400,426,429) and the square ldiv! (src/ql.jl:440-447,467).
//
// The dense work runs on the SAME FP64 tensor-core GEMMs as the real path.  For a complex matrix Z (a x b) write
//   [Z]   = its interleaved storage read as a real 2a x b matrix (rows re_0, im_0, re_1, im_1, ...), no copy needed;
//   [[Z]] = the real 2a x 2b matrix of 2x2 blocks [[re, -im], [im, re]].
// Then [Z Y] = [[Z]] [Y], [[Z Y]] = [[Z]] [[Y]] and [[Z^H]] = [[Z]]^T.  A complex block reflector H = I - V T V^H applied from
// the left is therefore the REAL block reflector I - [[V]] [[T]] [[V]]^T applied to [C]: only the thin operands V (p x ib) and T
// are expanded (kernels below), the big operand C is used in place, and the trailing update is exactly blocked.cu's
// apply_block with (2p, 2ib) -- three DMMA GEMMs with the same 8 real flops per complex multiply-add as a complex GEMM.
//
// The panel (64 complex columns) is factored column by column with ONE grid kernel per column (the end of step C -- zlarfg
// scalars + rank-1 update -- and the partial dot products of step C-1 on the same rows; every block re-sums the partials in
// the same fixed order -- bitwise reproducible), the complex analogue of the tall-strip kernels in panel.cu.  Norms are taken scaled by the column's largest entry and zlarfg's
// safmin loop is followed (LAPACK semantics for matrices scaled far away from 1).  Launch-bound for small n; the GEMMs
// dominate from n ~ 8192 on.  ComplexF32 is promoted to ComplexF64 (mixed.cu).
#include "common.cuh"
#include "internal.h"

#define QLB_REQUIRE(h, cond, argno, msg)                      \
    do {                                                      \
        if (!(cond)) return qlb_fail(h, -(argno), msg "%s");  \
    } while (0)

namespace qlb {

constexpr int ZT = 256;        // threads per block of the panel kernels
constexpr int ZNB = 64;        // complex columns per panel (128 real columns of [[V]]: the K = 128 update GEMMs of the real path)
constexpr int ZCH = 16;        // columns whose partial dot products a thread accumulates at a time
constexpr double Z_SAFMIN = 2.2250738585072014e-308 / 1.1102230246251565e-16;

__device__ __forceinline__ double2 zmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ double2 zconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double dlapy3_dev(double x, double y, double z) {
    const double xa = fabs(x), ya = fabs(y), za = fabs(z);
    const double w = fmax(xa, fmax(ya, za));
    if (w == 0.0 || w > 1.7976931348623157e308) return xa + ya + za;
    const double a = xa / w, b = ya / w, c = za / w;
    return w * sqrt(a * a + b * b + c * c);
}
// (1, 0) / (c, d), Smith's algorithm (LAPACK dladiv's classic form)
__device__ __forceinline__ double2 zrecip(double c, double d) {
    if (fabs(d) < fabs(c)) {
        const double e = d / c, f = c + d * e;
        return make_double2(1.0 / f, -e / f);
    }
    const double e = c / d, f = d + c * e;
    return make_double2(e / f, -1.0 / f);
}

// ---- the panel step ------------------------------------------------------------------------------------------------
// Per panel column C (pivot row mu, tail rows [0, mu)) ONE kernel does, for its block of rows,
//   (a) the end of step C: re-sum the partials of column C in a fixed order, zlarfg, scale the reflector, rank-1 update of
//       the panel columns j < C, clean reflector in block form
//         VV[2r, 2C] = re v_r, VV[2r+1, 2C] = im v_r, VV[2r, 2C+1] = -im v_r, VV[2r+1, 2C+1] = re v_r  (1 at the pivot, 0 below);
//   (b) the beginning of step C-1 on the rows it has just updated: partial sums for column C-1, the tail scaled by the
//       block's largest component amax_b:  part[b][j] = sum_i conj(x_i / amax_b) a_ij (j < C-1), part[b][C-1] = sum |x_i / amax_b|^2.
// (a) needs all blocks' partials of step C, which the previous launch wrote: one launch per column, two partial buffers
// alternate.  Layout of a partial buffer (doubles): nblk x ZNB x 2, then the pivot row ZNB x 2, then amax_b (nblk).
__device__ __forceinline__ void z_dots_phase(const double2 *__restrict__ A, int64_t lda, int mu, int C, int w, double *__restrict__ part,
                                             int rpb, double (*red)[ZCH][2], double *mx) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * rpb, r1 = min(r0 + rpb, mu);
    const double2 *xc = A + (int64_t)C * lda;
    double am = 0.0;
    for (int r = r0 + tid; r < r1; r += ZT) {
        const double2 x = xc[r];
        am = fmax(am, fmax(fabs(x.x), fabs(x.y)));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
    if (lane == 0) mx[warp] = am;
    __syncthreads();
    am = mx[0];
    for (int q = 1; q < ZT / 32; ++q) am = fmax(am, mx[q]);
    const double ram = am > 0.0 ? 1.0 / am : 0.0;
    // columns in chunks of ZCH (register accumulators); x is re-read per chunk
    for (int c0 = 0; c0 <= C; c0 += ZCH) {
        double dr[ZCH], di[ZCH];
#pragma unroll
        for (int j = 0; j < ZCH; ++j) dr[j] = di[j] = 0.0;
        if (am > 0.0) {
            for (int r = r0 + tid; r < r1; r += ZT) {
                double2 x = xc[r];
                x.x *= ram;
                x.y *= ram;
#pragma unroll
                for (int j = 0; j < ZCH; ++j) {
                    if (c0 + j < C) {
                        const double2 a = A[r + (int64_t)(c0 + j) * lda];
                        dr[j] = fma(x.x, a.x, fma(x.y, a.y, dr[j]));        // conj(x) a
                        di[j] = fma(x.x, a.y, fma(-x.y, a.x, di[j]));
                    } else if (c0 + j == C) {
                        dr[j] = fma(x.x, x.x, fma(x.y, x.y, dr[j]));
                    }
                }
            }
        }
        // warp totals of the 2 ZCH = 32 sums by a transposing reduction (31 shuffle pairs instead of 160): lane l ends up with
        // the total of value l (l < ZCH: real parts, l >= ZCH: imaginary parts)
        {
            double v[2 * ZCH];
#pragma unroll
            for (int j = 0; j < ZCH; ++j) {
                v[j] = dr[j];
                v[ZCH + j] = di[j];
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = lane & off;
#pragma unroll
                for (int k = 0; k < off; ++k) {
                    const double send = up ? v[k] : v[k + off], keep = up ? v[k + off] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            red[warp][lane & (ZCH - 1)][lane / ZCH] = v[0];
        }
        __syncthreads();
        if (tid < 2 * ZCH) {
            const int j = tid >> 1, c = tid & 1;
            double t = 0.0;
            for (int q = 0; q < ZT / 32; ++q) t += red[q][j][c];
            if (c0 + j < ZNB) part[((int64_t)blockIdx.x * ZNB + c0 + j) * 2 + c] = t;
        }
        __syncthreads();
    }
    if ((int)blockIdx.x == mu / rpb && tid < 2 * ZNB) {      // the block that owns (and has just updated) the pivot row publishes it
        const int j = tid >> 1, c = tid & 1;
        const double2 a = (j < w) ? A[mu + (int64_t)j * lda] : make_double2(0.0, 0.0);
        part[((int64_t)gridDim.x * ZNB + j) * 2 + c] = c ? a.y : a.x;
    }
    if (tid == 0) part[(int64_t)(gridDim.x + 1) * ZNB * 2 + blockIdx.x] = am;
}

__device__ __forceinline__ void z_update_phase(double2 *__restrict__ A, int64_t lda, int mu, int C, const double *__restrict__ part,
                                               double2 *__restrict__ tau_out, double *__restrict__ VV, int64_t ldvv, int p, int rpb,
                                               double2 *tot, double2 *rowv, double2 *coef, double *sc, double2 (*psum)[ZNB]) {
    const int tid = threadIdx.x;
    const int nblk = gridDim.x;
    const double *amaxs = part + (int64_t)(nblk + 1) * ZNB * 2;
    // amax over the blocks, then the partials re-summed in a fixed order: column j by the four threads (j, g), g = blocks b = g mod 4
    {
        double am = 0.0;
        for (int b = tid; b < nblk; b += ZT) am = fmax(am, amaxs[b]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
        if ((tid & 31) == 0) sc[8 + (tid >> 5)] = am;
        __syncthreads();
        am = sc[8];
        for (int q = 1; q < ZT / 32; ++q) am = fmax(am, sc[8 + q]);
        const int j = tid & (ZNB - 1), g = tid / ZNB;
        double tr = 0.0, ti = 0.0;
        if (am > 0.0 && j <= C) {
            const double ra = 1.0 / am;
#pragma unroll 4
            for (int b = g; b < nblk; b += ZT / ZNB) {
                const double wb = amaxs[b] * ra;                      // <= 1
                const double f = (j == C) ? wb * wb : wb;
                tr = fma(part[((int64_t)b * ZNB + j) * 2], f, tr);
                ti = fma(part[((int64_t)b * ZNB + j) * 2 + 1], f, ti);
            }
        }
        psum[g][j] = make_double2(tr, ti);
        __syncthreads();
        if (tid < ZNB) {
            const double2 a0 = psum[0][tid], a1 = psum[1][tid], a2 = psum[2][tid], a3 = psum[3][tid];
            tot[tid] = make_double2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
            rowv[tid] = make_double2(part[((int64_t)nblk * ZNB + tid) * 2], part[((int64_t)nblk * ZNB + tid) * 2 + 1]);
            if (tid == 0) sc[0] = am;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // LAPACK zlarfg: beta real, tau complex; H = I - tau v v^H, the factorization applies H^H (conj(tau))
        const double amax = sc[0];
        double alphr = rowv[C].x, alphi = rowv[C].y;
        double beta = alphr;
        double2 tauc = make_double2(0.0, 0.0), f = make_double2(0.0, 0.0);
        double betai = alphi;                                  // tau = 0 leaves alpha (complex) in place
        if (amax > 0.0 || alphi != 0.0) {
            double am_s = amax;
            double xnorm = am_s * sqrt(tot[C].x);
            double be = -copysign(dlapy3_dev(alphr, alphi, xnorm), alphr);
            int knt = 0;
            if (fabs(be) < Z_SAFMIN) {
                const double rsafmn = 1.0 / Z_SAFMIN;
                do {
                    ++knt;
                    am_s *= rsafmn;
                    be *= rsafmn;
                    alphr *= rsafmn;
                    alphi *= rsafmn;
                } while (fabs(be) < Z_SAFMIN && knt < 20);
                xnorm = am_s * sqrt(tot[C].x);
                be = -copysign(dlapy3_dev(alphr, alphi, xnorm), alphr);
            }
            tauc = make_double2((be - alphr) / be, -alphi / be);
            const double2 s = zrecip(alphr - be, alphi);      // 1 / (alpha - beta)
            f = make_double2(am_s * s.x, am_s * s.y);          // v_i = (x_i / amax) f
            for (int q = 0; q < knt; ++q) be *= Z_SAFMIN;
            beta = be;
            betai = 0.0;
        }
        sc[1] = f.x;
        sc[2] = f.y;
        sc[3] = tauc.x;
        sc[4] = tauc.y;
        sc[5] = beta;
        sc[6] = betai;
    }
    __syncthreads();
    const double amax = sc[0];
    const double2 f = make_double2(sc[1], sc[2]);
    if (tid < ZNB) {
        const double2 nct = make_double2(-sc[3], sc[4]);       // -conj(tau)
        double2 s = make_double2(0.0, 0.0);
        if (tid < C) s = zmul(nct, make_double2(fma(f.x, tot[tid].x, fma(f.y, tot[tid].y, rowv[tid].x)),        // conj(f) tot + rowv
                                                 fma(f.x, tot[tid].y, fma(-f.y, tot[tid].x, rowv[tid].y))));
        coef[tid] = s;
        if (blockIdx.x == 0) {
            if (tid < C) A[mu + (int64_t)tid * lda] = make_double2(rowv[tid].x + s.x, rowv[tid].y + s.y);      // v[mu] = 1
            if (tid == C) {
                *tau_out = make_double2(sc[3], sc[4]);
                A[mu + (int64_t)C * lda] = make_double2(sc[5], sc[6]);
            }
        }
    }
    __syncthreads();
    const double ram = amax > 0.0 ? 1.0 / amax : 0.0;
    const int r0 = blockIdx.x * rpb;
    double *v0 = VV + (int64_t)(2 * C) * ldvv, *v1 = v0 + ldvv;
    for (int r = r0 + tid; r < r0 + rpb && r < p; r += ZT) {
        double2 v = make_double2(0.0, 0.0);
        if (r < mu) {
            const double2 x = A[r + (int64_t)C * lda];
            v = zmul(make_double2(x.x * ram, x.y * ram), f);
            A[r + (int64_t)C * lda] = v;
            // chunks of 8 columns, the next chunk's loads issued before this chunk's stores (A may alias itself for the compiler)
            double2 cur[8], nxt[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < C) cur[q] = A[r + (int64_t)q * lda];
            for (int j0 = 0; j0 < C; j0 += 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (j0 + 8 + q < C) nxt[q] = A[r + (int64_t)(j0 + 8 + q) * lda];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (j0 + q < C) {
                        const double2 c = coef[j0 + q];
                        double2 a = cur[q];
                        a.x = fma(c.x, v.x, fma(-c.y, v.y, a.x));
                        a.y = fma(c.x, v.y, fma(c.y, v.x, a.y));
                        A[r + (int64_t)(j0 + q) * lda] = a;
                    }
#pragma unroll
                for (int q = 0; q < 8; ++q) cur[q] = nxt[q];
            }
        } else if (r == mu) {
            v = make_double2(1.0, 0.0);
        }
        v0[2 * r] = v.x;
        v0[2 * r + 1] = v.y;
        v1[2 * r] = -v.y;
        v1[2 * r + 1] = v.x;
    }
}

// UPDATE: finish column C (pivot row mu) from part_in; DOTS: start column C-1 (pivot row mu-1) into part_out.
template <bool UPDATE, bool DOTS>
__global__ void __launch_bounds__(ZT) zstep_kernel(double2 *__restrict__ A, int64_t lda, int mu, int C, int w,
                                                   const double *__restrict__ part_in, double *__restrict__ part_out,
                                                   double2 *__restrict__ tau_out, double *__restrict__ VV, int64_t ldvv, int p, int rpb) {
    __shared__ double red[ZT / 32][ZCH][2];
    __shared__ double mx[ZT / 32];
    __shared__ double2 tot[ZNB], rowv[ZNB], coef[ZNB];
    __shared__ double2 psum[ZT / ZNB][ZNB];
    __shared__ double sc[8 + ZT / 32];
    static_assert(ZT / ZNB == 4 && 2 * ZCH == 32, "re-summation layout");
    if constexpr (UPDATE) {
        z_update_phase(A, lda, mu, C, part_in, tau_out, VV, ldvv, p, rpb, tot, rowv, coef, sc, psum);
        __syncthreads();      // this block's rows of the panel are up to date (only this block reads them below)
    }
    if constexpr (DOTS) {
        const int Cn = UPDATE ? C - 1 : C, mun = UPDATE ? mu - 1 : mu;
        z_dots_phase(A, lda, mun, Cn, w, part_out, rpb, red, mx);
    }
}

// The same step with the block's rows of the panel staged in shared memory (one row per thread, ZS rows per block; for panels
// of at most 148 * ZS rows): ONE round trip to global memory per step (cp.async of the row, overlapped with the re-summation
// of the partials) instead of one per 8-column chunk of the update plus one per chunk of the dot products -- the step is a
// chain of L2 latencies, not of bandwidth or flops.
constexpr int ZS = 128;
template <bool UPDATE, bool DOTS>
__global__ void __launch_bounds__(ZS) zstep_smem_kernel(double2 *__restrict__ A, int64_t lda, int mu, int C, int w,
                                                        const double *__restrict__ part_in, double *__restrict__ part_out,
                                                        double2 *__restrict__ tau_out, double *__restrict__ VV, int64_t ldvv, int p) {
    extern __shared__ __align__(16) unsigned char zs_raw[];
    double2 *rb = reinterpret_cast<double2 *>(zs_raw);      // [ZNB][ZS]: column j of this block's rows
    __shared__ double red[ZS / 32][ZCH][2];
    __shared__ double2 tot[ZNB], rowv[ZNB], coef[ZNB];
    __shared__ double2 psum[ZS / ZNB][ZNB];
    __shared__ double sc[8 + ZS / 32];
    static_assert(ZS / ZNB == 2 && 2 * ZCH == 32, "re-summation layout");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x * ZS + tid;
    const int nblk = gridDim.x;
    const bool staged = r < (UPDATE ? mu : mu + 1);
    for (int j = 0; j <= C; ++j) cp_async16(&rb[j * ZS + tid], &A[(staged ? r : 0) + (int64_t)j * lda], staged);
    cp_async_commit();
    if constexpr (UPDATE) {
        const double *amaxs = part_in + (int64_t)(nblk + 1) * ZNB * 2;
        double am = 0.0;
        for (int b = tid; b < nblk; b += ZS) am = fmax(am, amaxs[b]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
        if (lane == 0) sc[8 + warp] = am;
        __syncthreads();
        am = sc[8];
        for (int q = 1; q < ZS / 32; ++q) am = fmax(am, sc[8 + q]);
        {
            const int j = tid & (ZNB - 1), g = tid / ZNB;
            double tr = 0.0, ti = 0.0;
            if (am > 0.0 && j <= C) {
                const double ra = 1.0 / am;
#pragma unroll 4
                for (int b = g; b < nblk; b += ZS / ZNB) {
                    const double wb = amaxs[b] * ra;
                    const double f = (j == C) ? wb * wb : wb;
                    tr = fma(part_in[((int64_t)b * ZNB + j) * 2], f, tr);
                    ti = fma(part_in[((int64_t)b * ZNB + j) * 2 + 1], f, ti);
                }
            }
            psum[g][j] = make_double2(tr, ti);
        }
        __syncthreads();
        if (tid < ZNB) {
            tot[tid] = make_double2(psum[0][tid].x + psum[1][tid].x, psum[0][tid].y + psum[1][tid].y);
            rowv[tid] = make_double2(part_in[((int64_t)nblk * ZNB + tid) * 2], part_in[((int64_t)nblk * ZNB + tid) * 2 + 1]);
        }
        __syncthreads();
        if (tid == 0) {      // LAPACK zlarfg (see z_update_phase)
            double alphr = rowv[C].x, alphi = rowv[C].y;
            double beta = alphr, betai = alphi;
            double2 tauc = make_double2(0.0, 0.0), f = make_double2(0.0, 0.0);
            if (am > 0.0 || alphi != 0.0) {
                double am_s = am;
                double xnorm = am_s * sqrt(tot[C].x);
                double be = -copysign(dlapy3_dev(alphr, alphi, xnorm), alphr);
                int knt = 0;
                if (fabs(be) < Z_SAFMIN) {
                    const double rsafmn = 1.0 / Z_SAFMIN;
                    do {
                        ++knt;
                        am_s *= rsafmn;
                        be *= rsafmn;
                        alphr *= rsafmn;
                        alphi *= rsafmn;
                    } while (fabs(be) < Z_SAFMIN && knt < 20);
                    xnorm = am_s * sqrt(tot[C].x);
                    be = -copysign(dlapy3_dev(alphr, alphi, xnorm), alphr);
                }
                tauc = make_double2((be - alphr) / be, -alphi / be);
                const double2 sr = zrecip(alphr - be, alphi);
                f = make_double2(am_s * sr.x, am_s * sr.y);
                for (int q = 0; q < knt; ++q) be *= Z_SAFMIN;
                beta = be;
                betai = 0.0;
            }
            sc[0] = am;
            sc[1] = f.x;
            sc[2] = f.y;
            sc[3] = tauc.x;
            sc[4] = tauc.y;
            sc[5] = beta;
            sc[6] = betai;
        }
        __syncthreads();
        const double2 f = make_double2(sc[1], sc[2]);
        if (tid < ZNB) {
            const double2 nct = make_double2(-sc[3], sc[4]);       // -conj(tau)
            double2 cf = make_double2(0.0, 0.0);
            if (tid < C) cf = zmul(nct, make_double2(fma(f.x, tot[tid].x, fma(f.y, tot[tid].y, rowv[tid].x)),
                                                      fma(f.x, tot[tid].y, fma(-f.y, tot[tid].x, rowv[tid].y))));
            coef[tid] = cf;
            if (blockIdx.x == 0) {
                if (tid < C) A[mu + (int64_t)tid * lda] = make_double2(rowv[tid].x + cf.x, rowv[tid].y + cf.y);
                if (tid == C) {
                    *tau_out = make_double2(sc[3], sc[4]);
                    A[mu + (int64_t)C * lda] = make_double2(sc[5], sc[6]);
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        const double ram = am > 0.0 ? 1.0 / am : 0.0;
        if (r < p) {
            double2 v = make_double2(0.0, 0.0);
            if (r < mu) {
                const double2 x = rb[C * ZS + tid];
                v = zmul(make_double2(x.x * ram, x.y * ram), f);
                A[r + (int64_t)C * lda] = v;
#pragma unroll 4
                for (int j = 0; j < C; ++j) {
                    const double2 c = coef[j];
                    double2 a = rb[j * ZS + tid];
                    a.x = fma(c.x, v.x, fma(-c.y, v.y, a.x));
                    a.y = fma(c.x, v.y, fma(c.y, v.x, a.y));
                    rb[j * ZS + tid] = a;
                    A[r + (int64_t)j * lda] = a;
                }
            } else if (r == mu) {
                v = make_double2(1.0, 0.0);
            }
            double *v0 = VV + (int64_t)(2 * C) * ldvv, *v1 = v0 + ldvv;
            v0[2 * r] = v.x;
            v0[2 * r + 1] = v.y;
            v1[2 * r] = -v.y;
            v1[2 * r + 1] = v.x;
        }
    } else {
        cp_async_wait<0>();
        __syncthreads();
    }
    if constexpr (DOTS) {
        // step Cn on the rows this thread has just updated (its own shared-memory entries only: no barrier needed before)
        const int Cn = UPDATE ? C - 1 : C, mun = UPDATE ? mu - 1 : mu;
        const bool live = r < mun;
        double2 x = live ? rb[Cn * ZS + tid] : make_double2(0.0, 0.0);
        double am = fmax(fabs(x.x), fabs(x.y));
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
        __syncthreads();      // (sc is reused)
        if (lane == 0) sc[8 + warp] = am;
        __syncthreads();
        am = sc[8];
        for (int q = 1; q < ZS / 32; ++q) am = fmax(am, sc[8 + q]);
        const double ram = am > 0.0 ? 1.0 / am : 0.0;
        x.x *= ram;
        x.y *= ram;
        for (int c0 = 0; c0 <= Cn; c0 += ZCH) {
            double v[2 * ZCH];
#pragma unroll
            for (int j = 0; j < ZCH; ++j) {
                double dr = 0.0, di = 0.0;
                if (c0 + j < Cn) {
                    const double2 a = rb[(c0 + j) * ZS + tid];
                    dr = fma(x.x, a.x, x.y * a.y);              // conj(x) a
                    di = fma(x.x, a.y, -x.y * a.x);
                } else if (c0 + j == Cn) {
                    dr = fma(x.x, x.x, x.y * x.y);
                }
                v[j] = dr;
                v[ZCH + j] = di;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const bool up = lane & off;
#pragma unroll
                for (int k = 0; k < off; ++k) {
                    const double send = up ? v[k] : v[k + off], keep = up ? v[k + off] : v[k];
                    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            red[warp][lane & (ZCH - 1)][lane / ZCH] = v[0];
            __syncthreads();
            if (tid < 2 * ZCH) {
                const int j = tid >> 1, c = tid & 1;
                double t = 0.0;
                for (int q = 0; q < ZS / 32; ++q) t += red[q][j][c];
                if (c0 + j < ZNB) part_out[((int64_t)blockIdx.x * ZNB + c0 + j) * 2 + c] = t;
            }
            __syncthreads();
        }
        if (r == mun) {      // the owner of the next pivot row publishes it
            for (int j = 0; j < ZNB; ++j) {
                const double2 a = (j <= Cn) ? rb[j * ZS + tid] : make_double2(0.0, 0.0);
                part_out[((int64_t)nblk * ZNB + j) * 2] = a.x;
                part_out[((int64_t)nblk * ZNB + j) * 2 + 1] = a.y;
            }
        }
        if (tid == 0) part_out[(int64_t)(nblk + 1) * ZNB * 2 + blockIdx.x] = am;
    }
}

// [[V]] (2p x 2ib, ldvv) from the packed complex factors F (column j has its pivot at row mu0 + j)
__global__ void zexpand_v_kernel(const double2 *__restrict__ F, int64_t ldf, int p, int ib, int mu0, double *__restrict__ VV,
                                 int64_t ldvv) {
    const int j = blockIdx.y;
    const int mu = mu0 + j;
    double *v0 = VV + (int64_t)(2 * j) * ldvv, *v1 = v0 + ldvv;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p; r += gridDim.x * blockDim.x) {
        const double2 v = (r < mu) ? F[r + (int64_t)j * ldf] : ((r == mu) ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0));
        v0[2 * r] = v.x;
        v0[2 * r + 1] = v.y;
        v1[2 * r] = -v.y;
        v1[2 * r + 1] = v.x;
    }
}

// [[T]] (2ib x 2ib, ldt) from GG = [[V^H V]] (2ib x 2ib, ld 2ib) and tau: zlarft backward/columnwise,
//   T(i,i) = tau_i,  T(i+1:, i) = T(i+1:, i+1:) * (-tau_i G(i+1:, i)),  G(l,i) = v_l^H v_i = GG[2l, 2i] + i GG[2l+1, 2i].
__global__ void __launch_bounds__(64) zlarft_kernel(const double *__restrict__ GG, const double2 *__restrict__ tau, int ib,
                                                    double *__restrict__ TT, int64_t ldt) {
    extern __shared__ __align__(16) unsigned char zl_smem[];
    double2 (*Tc)[ZNB + 1] = reinterpret_cast<double2 (*)[ZNB + 1]>(zl_smem);
    double2 *y = reinterpret_cast<double2 *>(zl_smem) + ZNB * (ZNB + 1);
    const int l = threadIdx.x;
    const int ldg = 2 * ib;
    for (int i = ib - 1; i >= 0; --i) {
        const double2 ti = tau[i];
        if (l > i && l < ib) {
            const double2 g = make_double2(GG[2 * l + (int64_t)(2 * i) * ldg], GG[2 * l + 1 + (int64_t)(2 * i) * ldg]);
            y[l] = zmul(make_double2(-ti.x, -ti.y), g);
        }
        __syncthreads();
        if (l > i && l < ib) {
            double2 s = make_double2(0.0, 0.0);
            for (int q = i + 1; q <= l; ++q) {
                const double2 t = zmul(Tc[l][q], y[q]);
                s.x += t.x;
                s.y += t.y;
            }
            Tc[l][i] = s;
        } else if (l == i) {
            Tc[i][i] = ti;
        }
        __syncthreads();
    }
    for (int idx = l; idx < ib * ib; idx += blockDim.x) {
        const int r = idx % ib, c = idx / ib;
        const double2 t = (r >= c) ? Tc[r][c] : make_double2(0.0, 0.0);
        TT[2 * r + (int64_t)(2 * c) * ldt] = t.x;
        TT[2 * r + 1 + (int64_t)(2 * c) * ldt] = t.y;
        TT[2 * r + (int64_t)(2 * c + 1) * ldt] = -t.y;
        TT[2 * r + 1 + (int64_t)(2 * c + 1) * ldt] = t.x;
    }
}

// out (cols x rows, ldo) = in (rows x cols, ldi)^H
__global__ void zctranspose_kernel(const double2 *__restrict__ in, int64_t ldi, int rows, int cols, double2 *__restrict__ out, int64_t ldo) {
    __shared__ double2 tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = bx + threadIdx.x, c = by + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[r + (int64_t)c * ldi];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = by + threadIdx.x, r = bx + j;
        if (r < rows && c < cols) out[c + (int64_t)r * ldo] = zconj(tile[threadIdx.x][j]);
    }
}

// forward substitution with the tb x tb diagonal block of L for one right-hand side per thread
__global__ void ztrsm_diag_kernel(const double2 *__restrict__ L, int64_t ldl, int tb, double2 *__restrict__ B, int64_t ldb, int nrhs,
                                  int i0, int32_t *__restrict__ dinfo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && dinfo) {
        for (int q = 0; q < tb; ++q) {
            const double2 d = L[q + (int64_t)q * ldl];
            if (d.x == 0.0 && d.y == 0.0) {
                atomicMin(dinfo, i0 + q + 1);
                break;
            }
        }
    }
    if (c >= nrhs) return;
    double2 *b = B + (int64_t)c * ldb;
    for (int q = 0; q < tb; ++q) {
        const double2 d = L[q + (int64_t)q * ldl];
        double2 x = b[q];
        {       // x / d (Smith)
            const double2 rd = zrecip(d.x, d.y);
            x = zmul(x, rd);
        }
        b[q] = x;
        for (int r = q + 1; r < tb; ++r) {
            const double2 t = zmul(L[r + (int64_t)q * ldl], x);
            b[r].x -= t.x;
            b[r].y -= t.y;
        }
    }
}
// B[i, :] -= L[i, 0:tb] X[0:tb, :] for the rows i below the block (one row per thread)
__global__ void ztrsm_update_kernel(const double2 *__restrict__ L, int64_t ldl, int tb, int rem, const double2 *__restrict__ X,
                                    double2 *__restrict__ B, int64_t ldb, int nrhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rem) return;
    for (int c = 0; c < nrhs; ++c) {
        double2 s = B[i + (int64_t)c * ldb];
        for (int q = 0; q < tb; ++q) {
            const double2 t = zmul(L[i + (int64_t)q * ldl], X[q + (int64_t)c * ldb]);
            s.x -= t.x;
            s.y -= t.y;
        }
        B[i + (int64_t)c * ldb] = s;
    }
}
__global__ void zset_int_kernel(int32_t *p, int32_t v) { *p = v; }

static int z_rows_per_block(int p) {
    int rpb = 256;
    while ((p + rpb - 1) / rpb > 148) rpb += 256;
    return rpb;
}
static size_t z_part_doubles(int nblk) { return (size_t)(nblk + 1) * ZNB * 2 + nblk + 64; }

// T of the block whose clean block-form reflectors are in ws.Vw
static int z_build_T(qlb200_ctx *h, cudaStream_t st, const QlbWorkspace &ws, int p, int ib, const double2 *tau) {
    int rc = qlb_block_gram(h, st, ws, 2 * p, 2 * ib);
    if (rc) return rc;
    constexpr size_t smem = (size_t)(ZNB * (ZNB + 1) + ZNB) * sizeof(double2);
    static_assert(ZNB <= 64, "zlarft_kernel: one thread per row");
    if (!h->zlarft_attr_done) {
        QLB_CUDA(h, cudaFuncSetAttribute(zlarft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->zlarft_attr_done = true;
    }
    zlarft_kernel<<<1, 64, smem, st>>>(ws.G, tau, ib, ws.T, ws.ldt);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

}  // namespace qlb
using namespace qlb;

// Factor one panel (complex columns [c0, c0+ib), rows [0, p)) on stream st: one launch per column; reflectors in block form
// -> VV (ldvv).  smem_rows: the block's rows staged in shared memory (fast, but one 128 KB block per SM: for a stream that
// owns the whole GPU); otherwise the global-memory variant (two blocks per SM: for the 16-SM panel partition).
static int z_factor_panel(qlb200_ctx *h, cudaStream_t st, double2 *A, int64_t lda, int p, int c0, int ib, double2 *tau_panel, double *VV,
                          int64_t ldvv, double *part, bool smem_rows) {
    double2 *Ap = A + (int64_t)c0 * lda;
    if (smem_rows && p <= 148 * ZS) {
        const int nb_ = (p + ZS - 1) / ZS;
        double *pb[2] = {part, part + z_part_doubles(nb_)};
        constexpr size_t smem = (size_t)ZNB * ZS * sizeof(double2);
        if (!h->zstep_attr_done) {
            QLB_CUDA(h, cudaFuncSetAttribute(zstep_smem_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QLB_CUDA(h, cudaFuncSetAttribute(zstep_smem_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            QLB_CUDA(h, cudaFuncSetAttribute(zstep_smem_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            h->zstep_attr_done = true;
        }
        zstep_smem_kernel<false, true><<<nb_, ZS, smem, st>>>(Ap, lda, p - 1, ib - 1, ib, nullptr, pb[0], nullptr, nullptr, 0, p);
        QLB_LAUNCH_CHECK(h);
        for (int s = 0; s < ib; ++s) {
            const int C = ib - 1 - s, mu = p - 1 - s;
            if (C > 0)
                zstep_smem_kernel<true, true><<<nb_, ZS, smem, st>>>(Ap, lda, mu, C, ib, pb[s & 1], pb[(s + 1) & 1], tau_panel + C, VV, ldvv, p);
            else
                zstep_smem_kernel<true, false><<<nb_, ZS, smem, st>>>(Ap, lda, mu, C, ib, pb[s & 1], nullptr, tau_panel + C, VV, ldvv, p);
            QLB_LAUNCH_CHECK(h);
        }
        return 0;
    }
    const int rpb = z_rows_per_block(p);
    const int nblk = (p + rpb - 1) / rpb;
    double *pb[2] = {part, part + z_part_doubles(nblk)};
    zstep_kernel<false, true><<<nblk, ZT, 0, st>>>(Ap, lda, p - 1, ib - 1, ib, nullptr, pb[0], nullptr, nullptr, 0, p, rpb);
    QLB_LAUNCH_CHECK(h);
    for (int s = 0; s < ib; ++s) {
        const int C = ib - 1 - s, mu = p - 1 - s;
        if (C > 0)
            zstep_kernel<true, true><<<nblk, ZT, 0, st>>>(Ap, lda, mu, C, ib, pb[s & 1], pb[(s + 1) & 1], tau_panel + C, VV, ldvv, p, rpb);
        else
            zstep_kernel<true, false><<<nblk, ZT, 0, st>>>(Ap, lda, mu, C, ib, pb[s & 1], nullptr, tau_panel + C, VV, ldvv, p, rpb);
        QLB_LAUNCH_CHECK(h);
    }
    return 0;
}

int qlb_zgeqlf_dev(qlb200_ctx *h, int64_t m64, int64_t n64, double *Ad, int64_t lda, double *taud) {
    const int m = (int)m64, n = (int)n64, k = m < n ? m : n;
    if (k == 0) return 0;
    double2 *A = reinterpret_cast<double2 *>(Ad);
    double2 *tau = reinterpret_cast<double2 *>(taud);
    cudaStream_t st = h->stream;
    QlbWorkspace ws;
    int rc = qlb_block_workspace(h, 2 * (int64_t)m, n, 2 * ZNB, &ws);
    if (rc) return rc;
    const int rpb = z_rows_per_block(m);
    rc = qlb_reserve(h, &h->tall, &h->tall_bytes, 2 * z_part_doubles(148 + (m + rpb - 1) / rpb) * sizeof(double));
    if (rc) return rc;
    double *part = (double *)h->tall;
    const int npanels = (k + ZNB - 1) / ZNB;
    auto P1 = [&](int j) { return k - j * ZNB; };                            // reflector range of panel j: [P0(j), P1(j))
    auto P0 = [&](int j) { const int i0 = k - (j + 1) * ZNB; return i0 > 0 ? i0 : 0; };
    // Look-ahead on the SM partitions (as ql! and ul!): panel j+1 is factored on the 16-SM partition (global-memory step kernel,
    // two blocks per SM) while the other partition applies panel j.  The reflectors / T of consecutive panels alternate
    // between the handle's workspace and a second set carved from h->zws.
    // Pays only where the panel (8 192 column steps x ~25 us on 16 SMs at n = 8192) is shorter than the update: n = 8192 218 -> 207 ms,
    // n = 4096 68 -> 80 ms; so from min(m,n) = 8192 on (option lookahead = 1 forces it: tests).
    const bool la = h->gc_ok && h->opt_lookahead != 0 && npanels > 2 && (h->opt_lookahead == 1 || k >= 8192);
    if (!la) {
        for (int j = 0; j < npanels; ++j) {
            const int i0 = P0(j), ib = P1(j) - i0;
            const int p = m - k + P1(j), c0 = n - k + i0;
            rc = z_factor_panel(h, st, A, lda, p, c0, ib, tau + i0, ws.Vw, ws.ldv, part, true);
            if (rc) return rc;
            if (c0 > 0) {
                rc = z_build_T(h, st, ws, p, ib, tau + i0);
                if (rc) return rc;
                // C <- H^H C = (I - V T^H V^H) C:  [[T]]^T = [[T^H]]  ->  apply_block with tt = false
                rc = qlb_block_apply(h, st, ws, true, false, ws.Vw, ws.ldv, ws.T, ws.ldt, 2 * p, 2 * ib, Ad, 2 * lda, c0);
                if (rc) return rc;
            }
        }
        return 0;
    }
    constexpr int NB2 = 2 * ZNB;
    const size_t gram_part = (size_t)160 * NB2 * NB2;
    rc = qlb_reserve(h, &h->zws, &h->zws_bytes, ((size_t)ws.ldv * NB2 + 2 * (size_t)NB2 * NB2 + gram_part + 64) * sizeof(double));
    if (rc) return rc;
    QlbWorkspace wp[2];                       // panel-stream views: own V, T, Gram and split-K scratch
    wp[0] = ws;
    wp[1] = ws;
    {
        double *q = (double *)h->zws;
        wp[1].Vw = q;  q += (size_t)ws.ldv * NB2;
        wp[1].T = q;   q += (size_t)NB2 * NB2;
        wp[0].G = wp[1].G = q;  q += (size_t)NB2 * NB2;
        wp[0].part = wp[1].part = q;
        wp[0].part_elems = wp[1].part_elems = (int64_t)gram_part;
    }
    cudaStream_t G = h->gc_gemm, P = h->gc_panel;
    cudaEvent_t evP = h->ev_a, evN = h->ev_b;
    QLB_CUDA(h, cudaEventRecord(h->ev_gc[0], st));
    QLB_CUDA(h, cudaStreamWaitEvent(G, h->ev_gc[0], 0));
    QLB_CUDA(h, cudaStreamWaitEvent(P, h->ev_gc[0], 0));
    auto factor = [&](int j) -> int {         // on P: panel j into set j & 1, and its T when there is something left of it
        const int i0 = P0(j), ib = P1(j) - i0;
        const int p = m - k + P1(j), c0 = n - k + i0;
        int r = z_factor_panel(h, P, A, lda, p, c0, ib, tau + i0, wp[j & 1].Vw, ws.ldv, part, false);
        if (!r && c0 > 0) r = z_build_T(h, P, wp[j & 1], p, ib, tau + i0);
        return r;
    };
    rc = factor(0);
    if (rc) return rc;
    QLB_CUDA(h, cudaEventRecord(evP, P));
    for (int j = 0; j < npanels; ++j) {
        const int i0 = P0(j), ib = P1(j) - i0;
        const int p = m - k + P1(j), c0 = n - k + i0;
        const QlbWorkspace &w = wp[j & 1];
        QLB_CUDA(h, cudaStreamWaitEvent(G, evP, 0));            // panel j (V, T) is complete
        if (j + 1 < npanels) {
            const int d0 = n - k + P0(j + 1);                   // next panel = columns [d0, c0)
            rc = qlb_block_apply(h, G, ws, true, false, w.Vw, ws.ldv, w.T, ws.ldt, 2 * p, 2 * ib, Ad + (int64_t)d0 * lda * 2, 2 * lda, c0 - d0);
            if (rc) return rc;
            QLB_CUDA(h, cudaEventRecord(evN, G));
            QLB_CUDA(h, cudaStreamWaitEvent(P, evN, 0));
            rc = factor(j + 1);
            if (rc) return rc;
            QLB_CUDA(h, cudaEventRecord(evP, P));
            if (d0 > 0) {
                rc = qlb_block_apply(h, G, ws, true, false, w.Vw, ws.ldv, w.T, ws.ldt, 2 * p, 2 * ib, Ad, 2 * lda, d0);
                if (rc) return rc;
            }
        } else if (c0 > 0) {
            rc = qlb_block_apply(h, G, ws, true, false, w.Vw, ws.ldv, w.T, ws.ldt, 2 * p, 2 * ib, Ad, 2 * lda, c0);
            if (rc) return rc;
        }
    }
    QLB_CUDA(h, cudaEventRecord(h->ev_gc[1], G));
    QLB_CUDA(h, cudaStreamWaitEvent(st, h->ev_gc[1], 0));
    return 0;
}

// side 'L': C (m x n) <- op(Q) C; side 'R': C <- C op(Q) through (C op(Q))^H = op(Q)^H C^H.  trans 'N' or 'C'.
int qlb_zunmql_dev(qlb200_ctx *h, char side, char trans, int64_t m64, int64_t n64, int64_t k64, const double *Ad, int64_t lda,
                   const double *taud, double *Cd, int64_t ldc) {
    const int m = (int)m64, n = (int)n64, k = (int)k64;
    if (m == 0 || n == 0 || k == 0) return 0;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    cudaStream_t st = h->stream;
    if (!left) {
        const int64_t ldt = (n + 1) & ~(int64_t)1;
        int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, (size_t)ldt * m * 2 * sizeof(double));
        if (rc) return rc;
        double2 *Ct = (double2 *)h->ws2;
        dim3 g1((m + 31) / 32, (n + 31) / 32), g2((n + 31) / 32, (m + 31) / 32), b(32, 8);
        zctranspose_kernel<<<g1, b, 0, st>>>(reinterpret_cast<const double2 *>(Cd), ldc, m, n, Ct, ldt);
        QLB_LAUNCH_CHECK(h);
        rc = qlb_zunmql_dev(h, 'L', notran ? 'C' : 'N', n, m, k, Ad, lda, taud, (double *)Ct, ldt);
        if (rc) return rc;
        zctranspose_kernel<<<g2, b, 0, st>>>(Ct, ldt, n, m, reinterpret_cast<double2 *>(Cd), ldc);
        QLB_LAUNCH_CHECK(h);
        return 0;
    }
    const double2 *A = reinterpret_cast<const double2 *>(Ad);
    const double2 *tau = reinterpret_cast<const double2 *>(taud);
    const int nq = m, other = n;
    QlbWorkspace ws;
    int rc = qlb_block_workspace(h, 2 * (int64_t)nq, other, 2 * ZNB, &ws);
    if (rc) return rc;
    const int nblk = (k + ZNB - 1) / ZNB;
    // Q = H_k ... H_1 (src/ql.jl:331 / :361): Q C applies H_1 first (leftmost block first), Q^H C the other way round
    for (int bi = 0; bi < nblk; ++bi) {
        const int j = notran ? nblk - 1 - bi : bi;
        const int i1 = k - j * ZNB;
        const int i = (i1 - ZNB > 0) ? i1 - ZNB : 0;
        const int ib = i1 - i;
        const int p = nq - k + i + ib;
        dim3 g((p + 255) / 256 > 64 ? 64 : (p + 255) / 256, ib);
        zexpand_v_kernel<<<g, 256, 0, st>>>(A + (int64_t)i * lda, lda, p, ib, nq - k + i, ws.Vw, ws.ldv);
        QLB_LAUNCH_CHECK(h);
        rc = z_build_T(h, st, ws, p, ib, tau + i);
        if (rc) return rc;
        rc = qlb_block_apply(h, st, ws, true, /*tt=*/notran, ws.Vw, ws.ldv, ws.T, ws.ldt, 2 * p, 2 * ib, Cd, 2 * ldc, other);
        if (rc) return rc;
    }
    return 0;
}

int qlb_zqlsolve_dev(qlb200_ctx *h, int64_t n64, int64_t nrhs64, const double *Ad, int64_t lda, const double *taud, double *Bd,
                     int64_t ldb, int32_t *dinfo) {
    const int n = (int)n64, nrhs = (int)nrhs64;
    if (n == 0 || nrhs == 0) return 0;
    int rc = qlb_zunmql_dev(h, 'L', 'C', n, nrhs, n, Ad, lda, taud, Bd, ldb);      // src/ql.jl:445
    if (rc) return rc;
    cudaStream_t st = h->stream;
    if (dinfo) {
        zset_int_kernel<<<1, 1, 0, st>>>(dinfo, 0x7fffffff);
        QLB_LAUNCH_CHECK(h);
    }
    const double2 *A = reinterpret_cast<const double2 *>(Ad);
    double2 *B = reinterpret_cast<double2 *>(Bd);
    constexpr int TB = 32;
    for (int i = 0; i < n; i += TB) {                                            // src/ql.jl:467
        const int tb = (n - i) < TB ? (n - i) : TB;
        ztrsm_diag_kernel<<<(nrhs + 63) / 64, 64, 0, st>>>(A + i + (int64_t)i * lda, lda, tb, B + i, ldb, nrhs, i, dinfo);
        QLB_LAUNCH_CHECK(h);
        const int rem = n - i - tb;
        if (rem > 0) {
            ztrsm_update_kernel<<<(rem + 127) / 128, 128, 0, st>>>(A + (i + tb) + (int64_t)i * lda, lda, tb, rem, B + i, B + i + tb, ldb, nrhs);
            QLB_LAUNCH_CHECK(h);
        }
    }
    return 0;
}

// rq!(A) for ComplexF64 (reference src/rq.jl:401 -> LAPACK.gerqf!): zgerqf(A) = (zgeqlf(A^H))^H with identical tau -- the RQ
// reflectors are stored conjugated in the rows (H(i) = I - tau v v^H, conj(v) in row i), which is exactly the conjugate
// transpose of the QL reflectors of A^H; R = L^H.
int qlb_zgerqf_dev(qlb200_ctx *h, int64_t m64, int64_t n64, double *Ad, int64_t lda, double *taud) {
    const int m = (int)m64, n = (int)n64;
    if (m == 0 || n == 0) return 0;
    cudaStream_t st = h->stream;
    const int64_t ldt = n;
    int rc = qlb_reserve(h, &h->ws2, &h->ws2_bytes, (size_t)ldt * m * 2 * sizeof(double));
    if (rc) return rc;
    double2 *At = (double2 *)h->ws2;
    dim3 b(32, 8), g1((m + 31) / 32, (n + 31) / 32), g2((n + 31) / 32, (m + 31) / 32);
    zctranspose_kernel<<<g1, b, 0, st>>>(reinterpret_cast<const double2 *>(Ad), lda, m, n, At, ldt);
    QLB_LAUNCH_CHECK(h);
    rc = qlb_zgeqlf_dev(h, n, m, (double *)At, ldt, taud);
    if (rc) return rc;
    zctranspose_kernel<<<g2, b, 0, st>>>(At, ldt, n, m, reinterpret_cast<double2 *>(Ad), lda);
    QLB_LAUNCH_CHECK(h);
    return 0;
}
// lmul!/rmul! with RQ's packed Q (reference src/rq.jl:130-137,183-202,234-291 -> LAPACK.ormrq! = zunmrq): A is k x nq, row i
// holds conj(v_i); Q_rq = H(1)^H ... H(k)^H = Q_ql^H for the QL reflectors A^H, so op(Q_rq) is the other op of Q_ql.
int qlb_zunmrq_dev(qlb200_ctx *h, char side, char trans, int64_t m64, int64_t n64, int64_t k64, const double *Ad, int64_t lda,
                   const double *taud, double *Cd, int64_t ldc) {
    const int m = (int)m64, n = (int)n64, k = (int)k64;
    if (m == 0 || n == 0 || k == 0) return 0;
    const bool left = (side == 'L' || side == 'l');
    const bool notran = (trans == 'N' || trans == 'n');
    const int nq = left ? m : n;
    const int64_t ldt = nq;
    int rc = qlb_reserve(h, &h->zws, &h->zws_bytes, (size_t)ldt * k * 2 * sizeof(double));
    if (rc) return rc;
    double2 *At = (double2 *)h->zws;                // nq x k: column i = v_i
    dim3 b(32, 8), g((k + 31) / 32, (nq + 31) / 32);
    zctranspose_kernel<<<g, b, 0, h->stream>>>(reinterpret_cast<const double2 *>(Ad), lda, k, nq, At, ldt);
    QLB_LAUNCH_CHECK(h);
    return qlb_zunmql_dev(h, side, notran ? 'C' : 'N', m, n, k, (const double *)At, ldt, taud, Cd, ldc);
}

// ---------------------------------------------------------------------------------------------
// extern "C" entry points (include/qlb200.h)
// ---------------------------------------------------------------------------------------------
static int zcheck_fact(qlb200_ctx *h, int64_t m, int64_t n, const void *A, int64_t lda, const void *tau) {
    if (!h) return -1;
    QLB_REQUIRE(h, m >= 0 && m < (1 << 29), 2, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 3, "n out of range");
    QLB_REQUIRE(h, A != nullptr || m * n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (m > 1 ? m : 1), 5, "lda < max(1,m)");
    QLB_REQUIRE(h, tau != nullptr || m * n == 0, 6, "tau is null");
    return 0;
}
static int zcheck_unmql(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                        const void *tau, const void *C, int64_t ldc) {
    if (!h) return -1;
    const bool left = side == 'L' || side == 'l', right = side == 'R' || side == 'r';
    QLB_REQUIRE(h, left || right, 2, "side must be 'L' or 'R'");
    QLB_REQUIRE(h, trans == 'N' || trans == 'n' || trans == 'C' || trans == 'c', 3, "trans must be 'N' or 'C'");
    QLB_REQUIRE(h, m >= 0 && m < (1 << 29), 4, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 5, "n out of range");
    const int64_t nq = left ? m : n;
    QLB_REQUIRE(h, k >= 0 && k <= nq, 6, "k must satisfy 0 <= k <= order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, A != nullptr || k == 0, 7, "A is null");
    QLB_REQUIRE(h, lda >= (nq > 1 ? nq : 1), 8, "lda < order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, tau != nullptr || k == 0, 9, "tau is null");
    QLB_REQUIRE(h, C != nullptr || m * n == 0, 10, "C is null");
    QLB_REQUIRE(h, ldc >= (m > 1 ? m : 1), 11, "ldc < max(1,m)");
    return 0;
}
static int zcheck_solve(qlb200_ctx *h, int64_t n, int64_t nrhs, const void *A, int64_t lda, const void *tau, const void *B, int64_t ldb) {
    if (!h) return -1;
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 2, "n out of range");
    QLB_REQUIRE(h, nrhs >= 0 && nrhs < (1 << 29), 3, "nrhs out of range");
    QLB_REQUIRE(h, A != nullptr || n == 0, 4, "A is null");
    QLB_REQUIRE(h, lda >= (n > 1 ? n : 1), 5, "lda < max(1,n)");
    QLB_REQUIRE(h, tau != nullptr || n == 0, 6, "tau is null");
    QLB_REQUIRE(h, B != nullptr || n * nrhs == 0, 7, "B is null");
    QLB_REQUIRE(h, ldb >= (n > 1 ? n : 1), 8, "ldb < max(1,n) (DimensionMismatch)");
    return 0;
}
// stage a column-major host matrix of 16-byte elements into one of the handle's device buffers (leading dimension m)
static int zstage_in(qlb200_ctx *h, void **buf, size_t *have, const void *A, int64_t m, int64_t n, int64_t lda, size_t extra, double **dA) {
    int rc = qlb_reserve(h, buf, have, ((size_t)(m > 0 ? m : 1) * (n > 0 ? n : 1) + extra) * 16);
    if (rc) return rc;
    *dA = (double *)*buf;
    if (m > 0 && n > 0)
        QLB_CUDA(h, cudaMemcpy2DAsync(*dA, (size_t)m * 16, A, (size_t)lda * 16, (size_t)m * 16, (size_t)n, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

extern "C" {

int qlb200_zgeqlf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau) {
    int rc = zcheck_fact(h, m, n, dA, lda, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_zgeqlf_dev(h, m, n, (double *)dA, lda, (double *)dtau);
}
int qlb200_zgeqlf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau) {
    int rc = zcheck_fact(h, m, n, A, lda, tau);
    if (rc) return rc;
    const int64_t k = m < n ? m : n;
    if (k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA;
    rc = zstage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, 0, &dA);
    if (rc) return rc;
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, (size_t)k * 16);
    if (rc) return rc;
    double *dtau = (double *)h->stage2;
    rc = qlb_zgeqlf_dev(h, m, n, dA, m, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpy2DAsync(A, (size_t)lda * 16, dA, (size_t)m * 16, (size_t)m * 16, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * 16, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int qlb200_zunmql_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc) {
    int rc = zcheck_unmql(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_zunmql_dev(h, side, trans, m, n, k, (const double *)dA, lda, (const double *)dtau, (double *)dC, ldc);
}
int qlb200_zunmql(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                  const void *tau, void *C, int64_t ldc) {
    int rc = zcheck_unmql(h, side, trans, m, n, k, A, lda, tau, C, ldc);
    if (rc) return rc;
    if (m == 0 || n == 0 || k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    const int64_t nq = (side == 'L' || side == 'l') ? m : n;
    double *dA, *dC;
    rc = zstage_in(h, &h->stage, &h->stage_bytes, A, nq, k, lda, 0, &dA);
    if (rc) return rc;
    rc = zstage_in(h, &h->stage2, &h->stage2_bytes, C, m, n, ldc, (size_t)k + 1, &dC);
    if (rc) return rc;
    double *dtau = dC + (size_t)m * n * 2;
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)k * 16, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_zunmql_dev(h, side, trans, m, n, k, dA, nq, dtau, dC, m);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpy2DAsync(C, (size_t)ldc * 16, dC, (size_t)m * 16, (size_t)m * 16, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int qlb200_zgerqf_dev(qlb200_handle h, int64_t m, int64_t n, void *dA, int64_t lda, void *dtau) {
    int rc = zcheck_fact(h, m, n, dA, lda, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_zgerqf_dev(h, m, n, (double *)dA, lda, (double *)dtau);
}
int qlb200_zgerqf(qlb200_handle h, int64_t m, int64_t n, void *A, int64_t lda, void *tau) {
    int rc = zcheck_fact(h, m, n, A, lda, tau);
    if (rc) return rc;
    const int64_t k = m < n ? m : n;
    if (k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA;
    rc = zstage_in(h, &h->stage, &h->stage_bytes, A, m, n, lda, 0, &dA);
    if (rc) return rc;
    rc = qlb_reserve(h, &h->stage2, &h->stage2_bytes, (size_t)k * 16);
    if (rc) return rc;
    double *dtau = (double *)h->stage2;
    rc = qlb_zgerqf_dev(h, m, n, dA, m, dtau);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpy2DAsync(A, (size_t)lda * 16, dA, (size_t)m * 16, (size_t)m * 16, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaMemcpyAsync(tau, dtau, (size_t)k * 16, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}
static int zcheck_unmrq(qlb200_ctx *h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                        const void *tau, const void *C, int64_t ldc) {
    if (!h) return -1;
    const bool left = side == 'L' || side == 'l', right = side == 'R' || side == 'r';
    QLB_REQUIRE(h, left || right, 2, "side must be 'L' or 'R'");
    QLB_REQUIRE(h, trans == 'N' || trans == 'n' || trans == 'C' || trans == 'c', 3, "trans must be 'N' or 'C'");
    QLB_REQUIRE(h, m >= 0 && m < (1 << 29), 4, "m out of range");
    QLB_REQUIRE(h, n >= 0 && n < (1 << 29), 5, "n out of range");
    const int64_t nq = left ? m : n;
    QLB_REQUIRE(h, k >= 0 && k <= nq, 6, "k must satisfy 0 <= k <= order of Q (DimensionMismatch)");
    QLB_REQUIRE(h, A != nullptr || k == 0, 7, "A is null");
    QLB_REQUIRE(h, lda >= (k > 1 ? k : 1), 8, "lda < max(1,k)");
    QLB_REQUIRE(h, tau != nullptr || k == 0, 9, "tau is null");
    QLB_REQUIRE(h, C != nullptr || m * n == 0, 10, "C is null");
    QLB_REQUIRE(h, ldc >= (m > 1 ? m : 1), 11, "ldc < max(1,m)");
    return 0;
}
int qlb200_zunmrq_dev(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *dA, int64_t lda,
                      const void *dtau, void *dC, int64_t ldc) {
    int rc = zcheck_unmrq(h, side, trans, m, n, k, dA, lda, dtau, dC, ldc);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_zunmrq_dev(h, side, trans, m, n, k, (const double *)dA, lda, (const double *)dtau, (double *)dC, ldc);
}
int qlb200_zunmrq(qlb200_handle h, char side, char trans, int64_t m, int64_t n, int64_t k, const void *A, int64_t lda,
                  const void *tau, void *C, int64_t ldc) {
    int rc = zcheck_unmrq(h, side, trans, m, n, k, A, lda, tau, C, ldc);
    if (rc) return rc;
    if (m == 0 || n == 0 || k == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    const int64_t nq = (side == 'L' || side == 'l') ? m : n;
    double *dA, *dC;
    rc = zstage_in(h, &h->stage, &h->stage_bytes, A, k, nq, lda, 0, &dA);
    if (rc) return rc;
    rc = zstage_in(h, &h->stage2, &h->stage2_bytes, C, m, n, ldc, (size_t)k + 1, &dC);
    if (rc) return rc;
    double *dtau = dC + (size_t)m * n * 2;
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)k * 16, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_zunmrq_dev(h, side, trans, m, n, k, dA, k, dtau, dC, m);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpy2DAsync(C, (size_t)ldc * 16, dC, (size_t)m * 16, (size_t)m * 16, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int qlb200_zqlsolve_dev(qlb200_handle h, int64_t n, int64_t nrhs, const void *dA, int64_t lda, const void *dtau, void *dB, int64_t ldb) {
    int rc = zcheck_solve(h, n, nrhs, dA, lda, dtau, dB, ldb);
    if (rc) return rc;
    QLB_CUDA(h, cudaSetDevice(h->device));
    return qlb_zqlsolve_dev(h, n, nrhs, (const double *)dA, lda, (const double *)dtau, (double *)dB, ldb, nullptr);
}
int qlb200_zqlsolve(qlb200_handle h, int64_t n, int64_t nrhs, const void *A, int64_t lda, const void *tau, void *B, int64_t ldb) {
    int rc = zcheck_solve(h, n, nrhs, A, lda, tau, B, ldb);
    if (rc) return rc;
    if (n == 0 || nrhs == 0) return 0;
    QLB_CUDA(h, cudaSetDevice(h->device));
    double *dA, *dB;
    rc = zstage_in(h, &h->stage, &h->stage_bytes, A, n, n, lda, 0, &dA);
    if (rc) return rc;
    rc = zstage_in(h, &h->stage2, &h->stage2_bytes, B, n, nrhs, ldb, (size_t)n + 1, &dB);
    if (rc) return rc;
    double *dtau = dB + (size_t)n * nrhs * 2;
    QLB_CUDA(h, cudaMemcpyAsync(dtau, tau, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream));
    rc = qlb_zqlsolve_dev(h, n, nrhs, dA, n, dtau, dB, n, h->dinfo);
    if (rc) return rc;
    QLB_CUDA(h, cudaMemcpy2DAsync(B, (size_t)ldb * 16, dB, (size_t)n * 16, (size_t)n * 16, (size_t)nrhs, cudaMemcpyDeviceToHost, h->stream));
    int32_t info = 0;
    QLB_CUDA(h, cudaMemcpyAsync(&info, h->dinfo, sizeof(info), cudaMemcpyDeviceToHost, h->stream));
    QLB_CUDA(h, cudaStreamSynchronize(h->stream));
    return info == 0x7fffffff ? 0 : (int)info;
}

}  // extern "C"
