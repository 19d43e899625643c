// batched_wy.cuh -- batched Float64 QL for 32 < n <= 64: matrix resident in shared memory, 8-column panel in registers,
// compact-WY trailing update on the FP64 tensor pipe (DMMA.8x8x4).  Included by batched.cu (uses its TMA / ?larfg helpers).
//
// Per-matrix semantics = ql! for BlasFloat (LAPACK dgeqlf conventions, reference src/ql.jl:72, spec src/ql.jl:56-68).
// A 64 x 64 Float64 matrix (32 KB) does not fit the register tile of geqlf_batched_kernel, so:
//   * one warp owns one matrix; the matrix lives in the warp's shared-memory image (column pitch LD = 8 mod 16 doubles);
//   * panels of 8 columns, from the right.  The panel is held row-distributed in registers: lane l owns rows l and l+32
//     of all 8 columns, so the reflector dot products are lane-local FMAs finished by ONE 8-value butterfly all-reduce
//     per step -- no shared-memory traffic inside the 8 steps of a panel;
//   * the step body is uniform (a real loop, 8 trips): the pivot column always sits at panel index 7; after each step the
//     columns rotate one index up and the finished reflector re-enters at index 0.  The butterfly therefore also reduces
//     v_r' v_c for the finished reflectors c > r, which is exactly the Gram row the compact-WY factor needs:
//         M = H_0 H_1 ... H_7 = I - V S V',  S upper triangular,  S[r][r] = tau_r,
//         S[r][c] = -tau_r sum_{k=r+1..c} (v_r' v_k) S[k][c]      (lane c keeps column c of S, rotating with the panel);
//   * trailing update C <- C - V S (V' C) per 8-column tile of C as three small DMMA products on fragments read straight
//     from the image:  W = C'V (K = rows),  W2 = -W S',  C' += W2 V'.  The C fragment loaded for W (16-byte loads, rows
//     2t, 2t+1 of column g) is the accumulator fragment of the last product: every tile of C is read and written once per
//     panel.  tools/proto_wy_batched.py is the NumPy model of this arithmetic (checked against LAPACK dgeqlf).
// Orders that are not a multiple of 8 use the next tile order: the matrix sits in the bottom-right corner (TMA loads from
// negative coordinates zero-fill the rest), padded steps are skipped, the corner goes out with plain stores.
// Matrices whose squares leave the plain range (or with an exactly zero tail) are flagged and redone by geql2_warp.
#pragma once

namespace qlb {

template <int NT>
struct WyCfg {
    static_assert(NT % 8 == 0 && NT >= 8 && NT <= 64, "tile order");
    static constexpr int LD = (NT % 16 == 8) ? NT : NT + 8;      // 2 adjacent columns' 64-byte row blocks -> different bank halves
    static constexpr int NP = NT / 8;                            // panels
    static constexpr int NSLOT = (NT + 31) / 32;                 // rows per lane
    static constexpr int IMG_BYTES = NT * LD * 8;
    static constexpr int S_OFF = IMG_BYTES;                      // 8 x 8 doubles: -S of the current panel
    static constexpr int TAU_OFF = S_OFF + 512;                  // NT doubles
    static constexpr int BAR_OFF = TAU_OFF + 512;
    static constexpr int WARP_BYTES = BAR_OFF + 128;
    static constexpr int WARPS_MAX = (232448 / WARP_BYTES) > 12 ? 12 : (232448 / WARP_BYTES);
};

// Trailing update of panel p (columns 8p..8p+7, NB = p+1 row blocks of 8) on the column tiles jt0 .. p-1.
// g = lane >> 2, t = lane & 3 (DMMA fragment coordinates, common.cuh).
template <int NT, int NB>
__device__ __forceinline__ void wy_update(double *img, double *sS, const double *taus8, int jt0, int lane, int jt1 = NB - 1,
                                          bool build_s = true) {
    constexpr int LD = WyCfg<NT>::LD;
    constexpr int p = NB - 1;
    const int g = lane >> 2, t = lane & 3;
    // B fragments of W = C'V:  V[8mb + 2t + s][g]   (one 16-byte load per row block)
    // B fragments of C' += W2 V':  V[8mb + g][2t + s]
    double2 Vf[NB];
    double VT[NB][2];
    const double *vcol = img + (8 * p + g) * LD + 2 * t;
    const double *vt0 = img + (8 * p + 2 * t) * LD + g;
#pragma unroll
    for (int mb = 0; mb < NB; ++mb) {
        Vf[mb] = *reinterpret_cast<const double2 *>(vcol + 8 * mb);
        VT[mb][0] = vt0[8 * mb];
        VT[mb][1] = vt0[LD + 8 * mb];
    }
    // the image holds the PACKED panel: make the diagonal block explicit (unit pivot, zeros below it)
    {
        const int r0 = 2 * t, r1 = 2 * t + 1;                 // row inside the block, column g
        Vf[p].x = (r0 < g) ? Vf[p].x : ((r0 == g) ? 1.0 : 0.0);
        Vf[p].y = (r1 < g) ? Vf[p].y : ((r1 == g) ? 1.0 : 0.0);
        VT[p][0] = (g < r0) ? VT[p][0] : ((g == r0) ? 1.0 : 0.0);      // row g, column 2t
        VT[p][1] = (g < r1) ? VT[p][1] : ((g == r1) ? 1.0 : 0.0);      // row g, column 2t+1
    }
    // compact-WY factor of the panel: M = H_0 ... H_7 = I - V S V', S upper triangular, S[r][r] = tau_r,
    // S[r][c] = -tau_r sum_{k=r+1..c} (v_r' v_k) S[k][c].  Gram matrix G = V'V on the tensor pipe (the fragments are already in
    // registers), then lane c runs the recurrence for column c of S (G read as shared-memory broadcasts); -S goes to the scratch.
    if (build_s) {
        double g0 = 0.0, g1 = 0.0, h0 = 0.0, h1 = 0.0;
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
            if (mb & 1) {
                dmma884(h0, h1, Vf[mb].x, Vf[mb].x);
                dmma884(h0, h1, Vf[mb].y, Vf[mb].y);
            } else {
                dmma884(g0, g1, Vf[mb].x, Vf[mb].x);
                dmma884(g0, g1, Vf[mb].y, Vf[mb].y);
            }
        }
        *reinterpret_cast<double2 *>(sS + g * 8 + 2 * t) = make_double2(g0 + h0, g1 + h1);      // G[g][2t], G[g][2t+1] (row-major)
        __syncwarp();
        double Sc[8];
        const int c = lane & 7;
#pragma unroll
        for (int r = 7; r >= 0; --r) {
            double acc = 0.0;
#pragma unroll
            for (int k = r + 1; k < 8; ++k) acc = fma(sS[r * 8 + k], Sc[k], acc);
            const double tr = taus8[r];
            Sc[r] = (r == c) ? tr : ((r < c) ? -tr * acc : 0.0);
        }
        __syncwarp();
        if (lane < 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) sS[lane * 8 + i] = -Sc[i];      // column-major -S
        }
        __syncwarp();
    }
    const double Sf0 = sS[(2 * t) * 8 + g], Sf1 = sS[(2 * t + 1) * 8 + g];       // -S[g][2t+s]
#pragma unroll 1
    for (int jt = jt0; jt < jt1; ++jt) {
        double2 acc[NB];
        double *ccol = img + (8 * jt + g) * LD + 2 * t;
        double wa0 = 0.0, wa1 = 0.0, wb0 = 0.0, wb1 = 0.0;     // two accumulation chains (DMMA dependent latency ~26 cycles)
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
            acc[mb] = *reinterpret_cast<const double2 *>(ccol + 8 * mb);
            if (mb & 1) {
                dmma884(wb0, wb1, acc[mb].x, Vf[mb].x);
                dmma884(wb0, wb1, acc[mb].y, Vf[mb].y);
            } else {
                dmma884(wa0, wa1, acc[mb].x, Vf[mb].x);
                dmma884(wa0, wa1, acc[mb].y, Vf[mb].y);
            }
        }
        const double w0 = wa0 + wb0, w1 = wa1 + wb1;           // W[j0+g][2t], W[j0+g][2t+1]
        double u0 = 0.0, u1 = 0.0;                             // W2 = -W S'
        dmma884(u0, u1, w0, Sf0);
        dmma884(u0, u1, w1, Sf1);
#pragma unroll
        for (int mb = 0; mb < NB; ++mb) {
            dmma884(acc[mb].x, acc[mb].y, u0, VT[mb][0]);
            dmma884(acc[mb].x, acc[mb].y, u1, VT[mb][1]);
            *reinterpret_cast<double2 *>(ccol + 8 * mb) = acc[mb];
        }
    }
}

// NSTEP Householder steps of a panel (pivot rows 8p + r, r = r_hi ... r_hi - NSTEP + 1) on the NL column registers P[8-NL .. 7].
// The pivot column always sits at index 7: after a step the finished reflector goes straight into the image (packed: tail
// scaled, beta at the pivot, the L entries below untouched), every column moves one index up and a zero column enters at the
// bottom (zero columns contribute exact zeros and need no masks).  Two instantiations per panel: NL = 8 for r = 7..4 and NL = 4
// for r = 3..0 (the live columns of steps 3..0 are exactly indices 4..7).
template <int NT, int NL, int NSLOT>
__device__ __forceinline__ void wy_steps(double (&P)[8][NSLOT], const int (&row)[NSLOT], double *pcol, double *taus, int p, int r_hi, int pad,
                                         int lane, bool &bad) {
    constexpr int LD = WyCfg<NT>::LD;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int B0 = 8 - NL;                        // first register column of this body
    constexpr int SH = (NL == 8) ? 2 : 3;             // the total of value i ends up in lanes (i << SH) ...
#pragma unroll 1
    for (int r = r_hi; r > r_hi - 4; --r) {
        const int R = 8 * p + r;                       // pivot row = pivot column
        const bool active = R > pad;                   // R == pad: the matrix's first column (identity, tau = 0); R < pad: padding
        const int lp = R & 31;
        double x7[NSLOT], xm[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            x7[s] = P[7][s];
            xm[s] = (row[s] < R) ? x7[s] : 0.0;
        }
        double d[NL];                                  // d[NL-1] = |x|^2, the others: x' P_i over the tail rows
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            d[i] = xm[0] * P[B0 + i][0];
            if constexpr (NSLOT == 2) d[i] = fma(xm[1], P[B0 + i][1], d[i]);
        }
        // transposing all-reduce over the 32 lanes (each stage halves the values a lane carries), then a broadcast: every lane
        // sees the same bits
        {
            double hsum;
            if constexpr (NL == 8) {
                const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
                double e[4], f[2];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double send = b4 ? d[k] : d[k + 4], keep = b4 ? d[k + 4] : d[k];
                    e[k] = keep + __shfl_xor_sync(FULL, send, 16);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const double send = b3 ? e[k] : e[k + 2], keep = b3 ? e[k + 2] : e[k];
                    f[k] = keep + __shfl_xor_sync(FULL, send, 8);
                }
                const double send = b2 ? f[0] : f[1], keep = b2 ? f[1] : f[0];
                hsum = keep + __shfl_xor_sync(FULL, send, 4);
            } else {
                const bool b4 = lane & 16, b3 = lane & 8;
                double f[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const double send = b4 ? d[k] : d[k + 2], keep = b4 ? d[k + 2] : d[k];
                    f[k] = keep + __shfl_xor_sync(FULL, send, 16);
                }
                const double send = b3 ? f[0] : f[1], keep = b3 ? f[1] : f[0];
                hsum = keep + __shfl_xor_sync(FULL, send, 8);
                hsum += __shfl_xor_sync(FULL, hsum, 4);
            }
            hsum += __shfl_xor_sync(FULL, hsum, 2);
            hsum += __shfl_xor_sync(FULL, hsum, 1);
#pragma unroll
            for (int i = 0; i < NL; ++i) d[i] = __shfl_sync(FULL, hsum, i << SH);
        }
        double aR[NL];                                 // row R of the panel (slot 0 holds the pivot rows)
#pragma unroll
        for (int i = 0; i < NL; ++i) aR[i] = __shfl_sync(FULL, P[B0 + i][0], lp);
        const double ssq = d[NL - 1], alpha = aR[NL - 1];
        const double tot = fma(alpha, alpha, ssq);
        bad = bad || (active && !in_plain_range(ssq, tot));
        const double rinv = nr_rsqrt(tot);
        const double nrm = tot * rinv;
        double beta = -copysign(nrm, alpha);
        double tauk = fma(fabs(alpha), rinv, 1.0);
        double scale = nr_rcp(alpha - beta);
        beta = active ? beta : alpha;
        tauk = active ? tauk : 0.0;
        scale = active ? scale : 0.0;
        if (lane == 0) taus[R] = tauk;
        double v[NSLOT], vf[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            v[s] = xm[s] * scale;
            vf[s] = (row[s] == R) ? 1.0 : v[s];
            // the finished column, packed
            const double outv = (row[s] < R) ? v[s] : ((row[s] == R) ? beta : x7[s]);
            if (NT % 32 == 0 || row[s] < NT) pcol[r * LD + row[s]] = outv;
        }
#pragma unroll
        for (int i = NL - 2; i >= 0; --i) {
            const double si = tauk * fma(d[i], scale, aR[i]);          // tau * v' a_i
#pragma unroll
            for (int s = 0; s < NSLOT; ++s) P[B0 + i + 1][s] = fma(-si, vf[s], P[B0 + i][s]);
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) P[B0][s] = 0.0;
    }
}

// TI = element type in global memory: double, or float (Float32 matrices are widened while they are staged -- plain loads,
// no TMA -- factored in Float64 and narrowed on the way out; flagged matrices are redone by the Float32 ?geql2 routine).
template <typename TI, int NT, int WARPS, bool USE_TMA, bool PADDED>
__global__ void __launch_bounds__(WARPS * 32, 1)
geqlf_wy_kernel(const __grid_constant__ CUtensorMap tmap, TI *__restrict__ A, int64_t lda, int64_t strideA,
                TI *__restrict__ tau, int64_t strideTau, int64_t batch, int nr_arg) {
    static_assert(!USE_TMA || sizeof(TI) == 8, "the tensor map describes Float64 matrices");
    using Cfg = WyCfg<NT>;
    constexpr int LD = Cfg::LD, NSLOT = Cfg::NSLOT, NP = Cfg::NP;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *wbase = smem_raw + (size_t)warp * Cfg::WARP_BYTES;
    double *img = reinterpret_cast<double *>(wbase);
    double *sS = reinterpret_cast<double *>(wbase + Cfg::S_OFF);
    double *taus = reinterpret_cast<double *>(wbase + Cfg::TAU_OFF);
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + Cfg::BAR_OFF);
    const int nr = PADDED ? nr_arg : NT;
    const int pad = NT - nr;
    const int nbatch = (int)batch;
    const int gw = (int)blockIdx.x * WARPS + warp, nw = (int)gridDim.x * WARPS;
    const bool flat4 = sizeof(TI) == 4 && lda == NT && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && ((strideA * sizeof(TI)) & 15) == 0;
    (void)flat4;
    if constexpr (USE_TMA) {
        if (lane == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncwarp();
    }
    uint32_t phase = 0;
    for (int mi = gw; mi < nbatch; mi += nw) {
        TI *Am = A + (int64_t)mi * strideA;
        if constexpr (USE_TMA) {
            if (lane == 0) {
                mbar_arrive_expect_tx(bar, (uint32_t)Cfg::IMG_BYTES);
                tma_load_3d(img, &tmap, -pad, -pad, mi, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            __syncwarp();
        } else {
            bool staged = false;
            if constexpr (sizeof(TI) == 4 && !PADDED) {
                if (flat4) {      // contiguous Float32 matrix: whole-warp 16-byte loads (4 rows of one column each)
#pragma unroll 8
                    for (int idx = lane * 4; idx < NT * NT; idx += 128) {
                        const float4 v = *reinterpret_cast<const float4 *>(Am + idx);
                        double *d = img + (idx / NT) * LD + (idx % NT);
                        *reinterpret_cast<double2 *>(d) = make_double2((double)v.x, (double)v.y);
                        *reinterpret_cast<double2 *>(d + 2) = make_double2((double)v.z, (double)v.w);
                    }
                    staged = true;
                }
            }
            if (!staged) {
                for (int c = 0; c < NT; ++c)
                    for (int r = lane; r < NT; r += 32) img[c * LD + r] = (c >= pad && r >= pad) ? (double)Am[(int64_t)(c - pad) * lda + (r - pad)] : 0.0;
            }
            __syncwarp();
        }
        bool bad = false;
#pragma unroll 1
        for (int p = NP - 1; p >= 0; --p) {
            if (8 * p + 7 < pad) break;                        // the remaining panels are padding
            // slot 0 = the 32-row half that holds this panel's pivot rows (8p .. 8p+7 never straddle a multiple of 32), slot 1 the other
            // half: the pivot-row broadcast needs no per-column select
            double P[8][NSLOT];
            int row[NSLOT];
            row[0] = lane + ((8 * p) & 32);
            if constexpr (NSLOT == 2) row[1] = lane + (((8 * p) & 32) ^ 32);
            const double *pcol = img + (8 * p) * LD;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int s = 0; s < NSLOT; ++s) P[i][s] = (NT % 32 == 0 || row[s] < NT) ? pcol[i * LD + row[s]] : 0.0;
            double *pw = img + (8 * p) * LD;
            wy_steps<NT, 8, NSLOT>(P, row, pw, taus, p, 7, pad, lane, bad);
            wy_steps<NT, 4, NSLOT>(P, row, pw, taus, p, 3, pad, lane, bad);
            __syncwarp();
            const int jt0 = pad >> 3;                          // column tiles left of it are padding
            if (jt0 < p) {
                switch (p) {
                    case 1: wy_update<NT, 2>(img, sS, taus + 8 * p, jt0, lane); break;
                    case 2: if constexpr (NT >= 24) wy_update<NT, 3>(img, sS, taus + 8 * p, jt0, lane); break;
                    case 3: if constexpr (NT >= 32) wy_update<NT, 4>(img, sS, taus + 8 * p, jt0, lane); break;
                    case 4: if constexpr (NT >= 40) wy_update<NT, 5>(img, sS, taus + 8 * p, jt0, lane); break;
                    case 5: if constexpr (NT >= 48) wy_update<NT, 6>(img, sS, taus + 8 * p, jt0, lane); break;
                    case 6: if constexpr (NT >= 56) wy_update<NT, 7>(img, sS, taus + 8 * p, jt0, lane); break;
                    case 7: if constexpr (NT >= 64) wy_update<NT, 8>(img, sS, taus + 8 * p, jt0, lane); break;
                    default: break;
                }
            }
            __syncwarp();
        }
        // ---- output ----
        if (!bad) {
            for (int i = lane; i < nr; i += 32) tau[(int64_t)mi * strideTau + i] = (TI)taus[i + pad];
        }
        if constexpr (USE_TMA && !PADDED) {
            fence_proxy_async();
            __syncwarp();
            if (!bad && lane == 0) tma_store_3d(&tmap, 0, 0, mi, img);
            tma_store_commit();
            tma_store_wait_read();
            __syncwarp();
        } else {
            if (!bad) {
                bool done = false;
                if constexpr (USE_TMA) {
                    if ((pad & 1) == 0) {
                        const int nv = nr >> 1;
                        for (int c = 0; c < nr; ++c)
                            for (int vv = lane; vv < nv; vv += 32)
                                *reinterpret_cast<double2 *>(Am + (int64_t)c * lda + 2 * vv) = *reinterpret_cast<const double2 *>(img + (c + pad) * LD + pad + 2 * vv);
                        done = true;
                    }
                }
                if constexpr (sizeof(TI) == 4 && !PADDED) {
                    if (flat4) {
#pragma unroll 8
                        for (int idx = lane * 4; idx < NT * NT; idx += 128) {
                            const double *d = img + (idx / NT) * LD + (idx % NT);
                            const double2 a0 = *reinterpret_cast<const double2 *>(d), a1 = *reinterpret_cast<const double2 *>(d + 2);
                            *reinterpret_cast<float4 *>(Am + idx) = make_float4((float)a0.x, (float)a0.y, (float)a1.x, (float)a1.y);
                        }
                        done = true;
                    }
                }
                if (!done) {
                    for (int c = 0; c < nr; ++c)
                        for (int r = lane; r < nr; r += 32) Am[(int64_t)c * lda + r] = (TI)img[(c + pad) * LD + pad + r];
                }
            }
            if constexpr (USE_TMA) fence_proxy_async();
            __syncwarp();
        }
        if (bad) {        // rare: the matrix is untouched in global memory; literal ?geql2 / ?larfg with the image as scratch
            geql2_warp<TI>(Am, lda, tau + (int64_t)mi * strideTau, nr, reinterpret_cast<TI *>(img), lane);
            if constexpr (USE_TMA) fence_proxy_async();
            __syncwarp();
        }
    }
    if constexpr (USE_TMA) tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// Warp-pair variant: TWO warps per matrix.  The panel warp runs the Householder steps, the update warp the tensor-core
// trailing updates, with a one-panel look-ahead inside the matrix: as soon as the update warp has applied panel p to the 8
// columns of panel p-1 it hands them over, and the panel warp factors panel p-1 while the update warp applies panel p to the
// rest.  The kernel is bound by the latency of the dependent step chain with only NT-limited matrices resident per SM
// (shared memory); this takes the update off that chain and doubles the warps per SM.  Hand-offs are mbarriers in shared
// memory (release/acquire at CTA scope), the pair meets on a named barrier around staging and output.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int NT>
__device__ __forceinline__ void wy_update_dispatch(int p, double *img, double *sS, const double *taus8, int jt0, int lane, int jt1, bool build_s) {
    switch (p) {
        case 1: wy_update<NT, 2>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        case 2: if constexpr (NT >= 24) wy_update<NT, 3>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        case 3: if constexpr (NT >= 32) wy_update<NT, 4>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        case 4: if constexpr (NT >= 40) wy_update<NT, 5>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        case 5: if constexpr (NT >= 48) wy_update<NT, 6>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        case 6: if constexpr (NT >= 56) wy_update<NT, 7>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        case 7: if constexpr (NT >= 64) wy_update<NT, 8>(img, sS, taus8, jt0, lane, jt1, build_s); break;
        default: break;
    }
}

template <typename TI, int NT, int PAIRS, bool USE_TMA, bool PADDED>
__global__ void __launch_bounds__(PAIRS * 64, 1)
geqlf_wy_pair_kernel(const __grid_constant__ CUtensorMap tmap, TI *__restrict__ A, int64_t lda, int64_t strideA,
                     TI *__restrict__ tau, int64_t strideTau, int64_t batch, int nr_arg) {
    static_assert(!USE_TMA || sizeof(TI) == 8, "the tensor map describes Float64 matrices");
    using Cfg = WyCfg<NT>;
    constexpr int LD = Cfg::LD, NSLOT = Cfg::NSLOT, NP = Cfg::NP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp >> 1, role = warp & 1;            // role 0: panel warp, role 1: update warp
    const int l64 = (role << 5) | lane;                     // index inside the pair
    unsigned char *wbase = smem_raw + (size_t)slot * Cfg::WARP_BYTES;
    double *img = reinterpret_cast<double *>(wbase);
    double *sS = reinterpret_cast<double *>(wbase + Cfg::S_OFF);
    double *taus = reinterpret_cast<double *>(wbase + Cfg::TAU_OFF);
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase + Cfg::BAR_OFF);      // [0] TMA, [1] "panel ready", [2] "next panel's columns updated"
    int *badflag = reinterpret_cast<int *>(wbase + Cfg::BAR_OFF + 32);
    const int nr = PADDED ? nr_arg : NT;
    const int pad = NT - nr;
    const int jt0 = pad >> 3;                               // column tiles left of it are padding
    const int nbatch = (int)batch;
    const int gp = (int)blockIdx.x * PAIRS + slot, npairs = (int)gridDim.x * PAIRS;
    const bool flat4 = sizeof(TI) == 4 && lda == NT && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && ((strideA * sizeof(TI)) & 15) == 0;
    (void)flat4;
    if (l64 == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], 1);
        mbar_fence_init();
    }
    pair_sync(1 + slot);
    uint32_t ph_tma = 0, ph_a = 0, ph_b = 0;
    for (int mi = gp; mi < nbatch; mi += npairs) {
        TI *Am = A + (int64_t)mi * strideA;
        if constexpr (USE_TMA) {
            if (l64 == 0) {
                mbar_arrive_expect_tx(&bar[0], (uint32_t)Cfg::IMG_BYTES);
                tma_load_3d(img, &tmap, -pad, -pad, mi, &bar[0]);
            }
            mbar_wait(&bar[0], ph_tma);
            ph_tma ^= 1;
        } else {
            bool staged = false;
            if constexpr (sizeof(TI) == 4 && !PADDED) {
                if (flat4) {
#pragma unroll 8
                    for (int idx = l64 * 4; idx < NT * NT; idx += 256) {
                        const float4 v = *reinterpret_cast<const float4 *>(Am + idx);
                        double *d = img + (idx / NT) * LD + (idx % NT);
                        *reinterpret_cast<double2 *>(d) = make_double2((double)v.x, (double)v.y);
                        *reinterpret_cast<double2 *>(d + 2) = make_double2((double)v.z, (double)v.w);
                    }
                    staged = true;
                }
            }
            if (!staged) {
                for (int c = role; c < NT; c += 2)
                    for (int r = lane; r < NT; r += 32) img[c * LD + r] = (c >= pad && r >= pad) ? (double)Am[(int64_t)(c - pad) * lda + (r - pad)] : 0.0;
            }
            pair_sync(1 + slot);
        }
        if (role == 0) {
            bool bad = false;
#pragma unroll 1
            for (int p = NP - 1; p >= 0; --p) {
                if (8 * p + 7 < pad) break;                        // the remaining panels are padding
                if (p + 1 < NP && jt0 < p + 1) {                   // panel p+1 has been applied to these 8 columns
                    mbar_wait(&bar[2], ph_b);
                    ph_b ^= 1;
                }
                double P[8][NSLOT];
                int row[NSLOT];
                row[0] = lane + ((8 * p) & 32);
                if constexpr (NSLOT == 2) row[1] = lane + (((8 * p) & 32) ^ 32);
                double *pw = img + (8 * p) * LD;
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) P[i][s] = (NT % 32 == 0 || row[s] < NT) ? pw[i * LD + row[s]] : 0.0;
                wy_steps<NT, 8, NSLOT>(P, row, pw, taus, p, 7, pad, lane, bad);
                wy_steps<NT, 4, NSLOT>(P, row, pw, taus, p, 3, pad, lane, bad);
                __syncwarp();
                if (p >= 1 && jt0 < p && lane == 0) mbar_arrive(&bar[1]);      // release: the panel's columns and tau are in shared memory
            }
            if (lane == 0) *badflag = bad ? 1 : 0;
        } else {
#pragma unroll 1
            for (int q = NP - 1; q >= 1; --q) {
                if (!(jt0 < q)) break;
                mbar_wait(&bar[1], ph_a);
                ph_a ^= 1;
                wy_update_dispatch<NT>(q, img, sS, taus + 8 * q, q - 1, lane, q, true);        // look-ahead: the next panel's columns first
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar[2]);
                if (jt0 < q - 1) wy_update_dispatch<NT>(q, img, sS, taus + 8 * q, jt0, lane, q - 1, false);
                __syncwarp();
            }
        }
        pair_sync(1 + slot);
        const bool bad = *badflag != 0;
        // ---- output ----
        if (!bad) {
            for (int i = l64; i < nr; i += 64) tau[(int64_t)mi * strideTau + i] = (TI)taus[i + pad];
        }
        if constexpr (USE_TMA && !PADDED) {
            fence_proxy_async();
            pair_sync(1 + slot);
            if (l64 == 0) {
                if (!bad) tma_store_3d(&tmap, 0, 0, mi, img);
                tma_store_commit();
                tma_store_wait_read();
            }
            pair_sync(1 + slot);
        } else {
            if (!bad) {
                bool done = false;
                if constexpr (USE_TMA) {
                    if ((pad & 1) == 0) {
                        const int nv = nr >> 1;
                        for (int c = role; c < nr; c += 2)
                            for (int vv = lane; vv < nv; vv += 32)
                                *reinterpret_cast<double2 *>(Am + (int64_t)c * lda + 2 * vv) = *reinterpret_cast<const double2 *>(img + (c + pad) * LD + pad + 2 * vv);
                        done = true;
                    }
                }
                if constexpr (sizeof(TI) == 4 && !PADDED) {
                    if (flat4) {
#pragma unroll 8
                        for (int idx = l64 * 4; idx < NT * NT; idx += 256) {
                            const double *d = img + (idx / NT) * LD + (idx % NT);
                            const double2 a0 = *reinterpret_cast<const double2 *>(d), a1 = *reinterpret_cast<const double2 *>(d + 2);
                            *reinterpret_cast<float4 *>(Am + idx) = make_float4((float)a0.x, (float)a0.y, (float)a1.x, (float)a1.y);
                        }
                        done = true;
                    }
                }
                if (!done) {
                    for (int c = role; c < nr; c += 2)
                        for (int r = lane; r < nr; r += 32) Am[(int64_t)c * lda + r] = (TI)img[(c + pad) * LD + pad + r];
                }
            }
            if constexpr (USE_TMA) fence_proxy_async();
            pair_sync(1 + slot);
        }
        if (bad) {        // rare: the matrix is untouched in global memory; literal ?geql2 / ?larfg with the image as scratch
            if (role == 0) geql2_warp<TI>(Am, lda, tau + (int64_t)mi * strideTau, nr, reinterpret_cast<TI *>(img), lane);
            if constexpr (USE_TMA) fence_proxy_async();
            pair_sync(1 + slot);
        }
    }
    if constexpr (USE_TMA) {
        if (l64 == 0) tma_store_wait_all();
    }
}

template <typename TI, int NT>
static int launch_geqlf_wy(qlb200_ctx *h, TI *A, int64_t lda, int64_t strideA, TI *tau, int64_t strideTau, int64_t batch, int nr) {
    using Cfg = WyCfg<NT>;
    constexpr int WARPS = Cfg::WARPS_MAX;
    static_assert(Cfg::IMG_BYTES + 1024 >= ((NT + 1) * NT + NT) * 8, "the image + scratch hold the generic routine's scratch");
    const size_t smem = (size_t)WARPS * Cfg::WARP_BYTES;
    if (batch >= ((int64_t)1 << 31)) return qlb_fail(h, -8, "batch too large%s");
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    void (*kern)(const CUtensorMap, TI *, int64_t, int64_t, TI *, int64_t, int64_t, int);
    if constexpr (sizeof(TI) == 8) {
        const bool tma = tma_ok(A, lda, strideA) && make_batch_tensor_map<double>(&tmap, A, nr, lda, strideA, batch, Cfg::LD, NT);
        if (nr == NT) kern = tma ? geqlf_wy_kernel<TI, NT, WARPS, true, false> : geqlf_wy_kernel<TI, NT, WARPS, false, false>;
        else kern = tma ? geqlf_wy_kernel<TI, NT, WARPS, true, true> : geqlf_wy_kernel<TI, NT, WARPS, false, true>;
    } else {
        kern = (nr == NT) ? geqlf_wy_kernel<TI, NT, WARPS, false, false> : geqlf_wy_kernel<TI, NT, WARPS, false, true>;
    }
    // The warp-pair kernel (two warps per matrix, look-ahead inside the matrix) wins where shared memory allows at most ~7
    // matrices per SM anyway (n > 56: 3.01 vs 3.32 ms per 65 536 64 x 64 matrices); below, one warp per matrix keeps more
    // matrices resident and wins (n = 48: 1.61 vs 2.06 ms, n = 40: 1.05 vs 1.66 ms).  bvar 6 / 7 force pair / single (tests).
    constexpr int PAIRS = WARPS < 7 ? WARPS : 7;
    int threads = WARPS * 32, per_cta = WARPS;
    size_t smem_used = smem;
    if (h->opt_bvar == 6 || (h->opt_bvar != 7 && NT >= 64)) {
        if constexpr (sizeof(TI) == 8) {
            const bool tma = tma_ok(A, lda, strideA) && make_batch_tensor_map<double>(&tmap, A, nr, lda, strideA, batch, Cfg::LD, NT);
            if (nr == NT) kern = tma ? geqlf_wy_pair_kernel<TI, NT, PAIRS, true, false> : geqlf_wy_pair_kernel<TI, NT, PAIRS, false, false>;
            else kern = tma ? geqlf_wy_pair_kernel<TI, NT, PAIRS, true, true> : geqlf_wy_pair_kernel<TI, NT, PAIRS, false, true>;
        } else {
            kern = (nr == NT) ? geqlf_wy_pair_kernel<TI, NT, PAIRS, false, false> : geqlf_wy_pair_kernel<TI, NT, PAIRS, false, true>;
        }
        threads = PAIRS * 64;
        per_cta = PAIRS;
        smem_used = (size_t)PAIRS * Cfg::WARP_BYTES;
    }
    int occ = 0;
    int rcs = kernel_ready(h, kern, threads, smem_used, &occ);
    if (rcs) return rcs;
    if (occ < 1) return qlb_fail(h, 1000 + (int)cudaErrorLaunchOutOfResources, "batched WY QL kernel does not fit on an SM%s");
    int64_t blocks = (batch + per_cta - 1) / per_cta;
    int64_t cap = (int64_t)h->sm_count * occ;
    if (h->opt_bgrid > 0 && cap > h->opt_bgrid) cap = h->opt_bgrid;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, threads, smem_used, h->stream>>>(tmap, A, lda, strideA, tau, strideTau, batch, nr);
    QLB_LAUNCH_CHECK(h);
    return 0;
}

}  // namespace qlb
