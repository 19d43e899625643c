/*
 * ql_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).  Not part of the product path.
 *
 * Plain-C restatement of the reference's Householder-QL hot path, type-generic through the
 * macros  T  (element type) and  FN(name)  (symbol suffixing).  Included twice by ql_oracle.c
 * (double -> orc_d*, float -> orc_s*).
 *
 * Every function cites the reference lines (relative to /root/reference) it restates, or -- for
 * the arithmetic that lives in a third-party dependency absent from /root/reference (LAPACK
 * inside OpenBLAS, reached from src/ql.jl:72, src/rq.jl:401; LinearAlgebra.reflector! reached
 * through ArrayLayouts, src/MatrixFactorizations.jl:29) -- the published algorithm it restates.
 *
 * Column-major, 0-based C indexing of 1-based reference loops: A(i,j) = A[i + j*lda].
 * Parity status: PINNED (tests/test_oracle.py) against
 *   - the LAPACK routine the reference calls (scipy-openblas dgeqlf/sgeqlf/dgerqf/dormql ...),
 *   - the reference's own relational known-answer tests (test/test_ql.jl:5-33 flip-QR identity,
 *     :120-124 Issue 7304 sign test, :151-155 wide QL, :157-181 lmul!/rmul! vs Matrix(Q)),
 *   - committed golden fixtures the .npz files under tests/golden/.
 */

#define A_(i, j) A[(size_t)(i) + (size_t)(j) * (size_t)lda]
#define B_(i, j) B[(size_t)(i) + (size_t)(j) * (size_t)ldb]

/* ---- dnrm2: scaled sum of squares (reference-LAPACK/BLAS dnrm2, classic scale/ssq form) ---- */
static T FN(nrm2)(int64_t n, const T *x, int64_t incx) {
    if (n < 1) return (T)0;
    if (n == 1) return FABS(x[0]);
    T scale = (T)0, ssq = (T)1;
    for (int64_t i = 0; i < n; ++i) {
        T v = x[i * incx];
        if (v != (T)0) {
            T a = FABS(v);
            if (scale < a) { T r = scale / a; ssq = (T)1 + ssq * r * r; scale = a; }
            else           { T r = a / scale; ssq += r * r; }
        }
    }
    return scale * SQRT(ssq);
}

/* dlapy2: sqrt(x^2+y^2) avoiding unnecessary overflow (reference LAPACK dlapy2). */
static T FN(lapy2)(T x, T y) {
    T xa = FABS(x), ya = FABS(y);
    T w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
    if (z == (T)0 || w > HUGEV) return w;
    T r = z / w;
    return w * SQRT((T)1 + r * r);
}

/*
 * ?larfg -- elementary reflector, LAPACK convention (what src/ql.jl:72 LAPACK.geqlf! executes
 * through ?geql2).  H = I - tau*[1;v][1;v]^T, H*[alpha;x] = [beta;0].
 *   tau = 0, beta = alpha when ||x|| == 0 (no sign flip)  <- differs from Julia's reflector!.
 */
void FN(larfg)(int64_t n, T *alpha, T *x, int64_t incx, T *tau) {
    if (n <= 1) { *tau = (T)0; return; }
    T xnorm = FN(nrm2)(n - 1, x, incx);
    if (xnorm == (T)0) { *tau = (T)0; return; }
    T beta = -COPYSIGN(FN(lapy2)(*alpha, xnorm), *alpha);
    const T safmin = SAFMIN / EPSV;
    int knt = 0;
    if (FABS(beta) < safmin) {
        const T rsafmn = (T)1 / safmin;
        do {
            ++knt;
            for (int64_t i = 0; i < n - 1; ++i) x[i * incx] *= rsafmn;
            beta *= rsafmn;
            *alpha *= rsafmn;
        } while (FABS(beta) < safmin && knt < 20);
        xnorm = FN(nrm2)(n - 1, x, incx);
        beta = -COPYSIGN(FN(lapy2)(*alpha, xnorm), *alpha);
    }
    *tau = (beta - *alpha) / beta;
    T sc = (T)1 / (*alpha - beta);
    for (int64_t i = 0; i < n - 1; ++i) x[i * incx] *= sc;
    for (int j = 0; j < knt; ++j) beta *= safmin;
    *alpha = beta;
}

/*
 * Julia LinearAlgebra.reflector!(x) (imported at src/MatrixFactorizations.jl:29, used at
 * src/ql.jl:63): x[0] is the pivot.  nu = copysign(||x||, x0); x0 <- -nu; x[1:] /= (x0+nu);
 * tau = (x0+nu)/nu; returns 0 for the all-zero vector.  No tau=0 shortcut for a zero TAIL:
 * x=[1,0,0] gives tau=2 and flips the sign.
 */
T FN(jl_reflector)(int64_t n, T *x, int64_t incx) {
    if (n == 0) return (T)0;
    T xi1 = x[0];
    T normu = FN(nrm2)(n, x, incx);
    if (normu == (T)0) return (T)0;
    T nu = COPYSIGN(normu, xi1);
    xi1 += nu;
    x[0] = -nu;
    for (int64_t i = 1; i < n; ++i) x[i * incx] /= xi1;
    return xi1 / nu;
}

/*
 * QL factorization, unblocked.  Restates generic_qlfactUnblocked! (src/ql.jl:56-68):
 *   for k = n:-1:max(n-m+1+(T<:Real),1): mu = m+k-n; x = A[mu:-1:1,k]; tau_k = reflector!(x);
 *   reflectorApply!(x, tau_k, A[mu:-1:1, 1:k-1])
 * convention = 0: LAPACK ?larfg conventions (= ?geql2, the arithmetic of the Float64/Float32
 *                 production path src/ql.jl:72; loop runs down to k = max(n-m+1,1), the final
 *                 length-1 reflector yields tau = 0 exactly as the (T<:Real) skip does).
 * convention = 1: Julia generic path (reflector! semantics, (T<:Real) loop bound).
 * tau has min(m,n) entries; tau index k-n+min(m,n) (src/ql.jl:64).
 */
void FN(geql2)(int64_t m, int64_t n, T *A, int64_t lda, T *tau, int convention) {
    int64_t mn = m < n ? m : n;
    for (int64_t i = 0; i < mn; ++i) tau[i] = (T)0;
    int64_t klo = n - m + 1 + (convention == 1 ? 1 : 0);       /* 1-based */
    if (klo < 1) klo = 1;
    for (int64_t k = n; k >= klo; --k) {                         /* 1-based column */
        int64_t mu = m + k - n;                                  /* 1-based pivot row */
        T *col = &A_(0, k - 1);
        T tk;
        if (convention == 0) {
            /* pivot = col[mu-1], tail = col[0..mu-2]  (order of the tail is irrelevant) */
            FN(larfg)(mu, &col[mu - 1], col, 1, &tk);
        } else {
            /* reversed view x = A[mu:-1:1,k]: x[0] = col[mu-1], stride -1 */
            tk = FN(jl_reflector)(mu, &col[mu - 1], -1);
        }
        tau[k - n + mn - 1] = tk;
        /* reflectorApply!(x, tau, A[mu:-1:1, 1:k-1]) : A <- (I - tau v v^T) A, v = [1; tail] */
        for (int64_t j = 0; j < k - 1; ++j) {
            T *cj = &A_(0, j);
            T vAj = cj[mu - 1];
            for (int64_t i = mu - 2; i >= 0; --i) vAj += col[i] * cj[i];
            vAj = tk * vAj;
            cj[mu - 1] -= vAj;
            for (int64_t i = mu - 2; i >= 0; --i) cj[i] -= col[i] * vAj;
        }
    }
}

/* getL (src/ql.jl:214-217): copy factors[end-min(m,n)+1:end, 1:n] then tril!(., max(n-m,0)).
 * L is min(m,n) x n, ldl >= min(m,n). */
void FN(getL)(int64_t m, int64_t n, const T *A, int64_t lda, T *L, int64_t ldl) {
    int64_t mn = m < n ? m : n, r0 = m - mn, kd = n > m ? n - m : 0;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < mn; ++i)
            L[i + j * ldl] = (j - i <= kd) ? A_(r0 + i, j) : (T)0;
}

/* lmul!(Q,B) (src/ql.jl:321-347, k ascending) / lmul!(Q',B) (src/ql.jl:350-377, k descending).
 * A = factors (mA x nA), B is mA x nB in place.  Returns -1 on DimensionMismatch (mA != mB). */
int FN(lmul_q)(int64_t mA, int64_t nA, const T *A, int64_t lda, const T *tau,
               int64_t mB, int64_t nB, T *B, int64_t ldb, int adjoint) {
    if (mA != mB) return -1;
    int64_t mn = mA < nA ? mA : nA;
    int64_t klo = nA - mA + 1 > 1 ? nA - mA + 1 : 1;
    for (int64_t kk = 0; kk <= nA - klo; ++kk) {
        int64_t k = adjoint ? nA - kk : klo + kk;               /* 1-based */
        int64_t mu = mA + k - nA;
        T tk = tau[k - nA + mn - 1];
        for (int64_t j = 0; j < nB; ++j) {
            T vBj = B_(mu - 1, j);
            for (int64_t i = 0; i < mu - 1; ++i) vBj += A_(i, k - 1) * B_(i, j);
            vBj = tk * vBj;
            B_(mu - 1, j) -= vBj;
            for (int64_t i = 0; i < mu - 1; ++i) B_(i, j) -= A_(i, k - 1) * vBj;
        }
    }
    return 0;
}

/* rmul!(A,Q) (src/ql.jl:381-406, k descending) / rmul!(A,Q') (src/ql.jl:409-435, k ascending).
 * Q factors are mQ x nQ; C is mC x mQ in place (named C here; the reference calls it A). */
int FN(rmul_q)(int64_t mQ, int64_t nQ, const T *A, int64_t lda, const T *tau,
               int64_t mC, int64_t nC, T *B, int64_t ldb, int adjoint) {
    if (nC != mQ) return -1;
    int64_t mn = mQ < nQ ? mQ : nQ;
    int64_t klo = nQ - mQ + 1 > 1 ? nQ - mQ + 1 : 1;
    for (int64_t kk = 0; kk <= nQ - klo; ++kk) {
        int64_t k = adjoint ? klo + kk : nQ - kk;
        int64_t mu = mQ + k - nQ;
        T tk = tau[k - nQ + mn - 1];
        for (int64_t i = 0; i < mC; ++i) {
            T vAi = B_(i, mu - 1);
            for (int64_t j = 0; j < mu - 1; ++j) vAi += B_(i, j) * A_(j, k - 1);
            vAi = vAi * tk;
            B_(i, mu - 1) -= vAi;
            for (int64_t j = 0; j < mu - 1; ++j) B_(i, j) -= vAi * A_(j, k - 1);
        }
    }
    return 0;
}

/* Matrix(Q) (src/ql.jl:264-266): lmul!(Q, I(m, min(m,n))).  Q out is m x min(m,n), ldq >= m. */
void FN(matrix_q)(int64_t m, int64_t n, const T *A, int64_t lda, const T *tau, T *Q, int64_t ldq) {
    int64_t mn = m < n ? m : n;
    for (int64_t j = 0; j < mn; ++j)
        for (int64_t i = 0; i < m; ++i) Q[i + j * ldq] = (i == j) ? (T)1 : (T)0;
    FN(lmul_q)(m, n, A, lda, tau, m, mn, Q, ldq, 0);
}

/* ldiv!(F::QL, B), square branch only (src/ql.jl:440-447,467): B <- L^{-1} (Q' B).
 * The tall/wide branches are @test_broken in the reference (test/test_ql.jl:148) and are not
 * restated.  Forward substitution stands in for the BLAS trsm ArrayLayouts dispatches to.
 * Returns -1 if not square / dimension mismatch, k+1 if L[k,k] == 0. */
int FN(ldiv_ql)(int64_t n, const T *A, int64_t lda, const T *tau, int64_t nB, T *B, int64_t ldb) {
    FN(lmul_q)(n, n, A, lda, tau, n, nB, B, ldb, 1);
    for (int64_t j = 0; j < nB; ++j)
        for (int64_t i = 0; i < n; ++i) {
            T s = B_(i, j);
            for (int64_t c = 0; c < i; ++c) s -= A_(i, c) * B_(c, j);
            if (A_(i, i) == (T)0) return (int)(i + 1);
            B_(i, j) = s / A_(i, i);
        }
    return 0;
}

/*
 * RQ factorization, unblocked.  convention = 0: LAPACK ?gerq2 (arithmetic of src/rq.jl:401
 * LAPACK.gerqf!): for i = k..1: larfg on row m-k+i (pivot A(m-k+i, n-k+i), tail A(m-k+i,1:n-k+i-1)),
 * then apply H(i) from the right to A(1:m-k+i-1, 1:n-k+i).
 * convention = 1: _rqfactUnblocked! + reflectorYlppa! (src/rq.jl:327-380) with reflector! semantics.
 */
void FN(gerq2)(int64_t m, int64_t n, T *A, int64_t lda, T *tau, int convention) {
    int64_t mm = m < n ? m : n;
    for (int64_t i = 0; i < mm; ++i) tau[i] = (T)0;
    for (int64_t k = mm; k >= 1; --k) {
        int64_t krow = m - mm + k, kcol = n - mm + k;           /* 1-based */
        T *row = &A_(krow - 1, 0);                              /* stride lda */
        T tk;
        if (convention == 0) {
            FN(larfg)(kcol, &row[(kcol - 1) * lda], row, lda, &tk);
        } else if (kcol == 1) {
            tk = (T)0;                                          /* src/rq.jl:366-367 */
        } else {
            /* v = [A[krow,kcol]; A[krow,1:kcol-1]] (src/rq.jl:362-365); reflector!(v) */
            T nrm;
            {   /* norm over the whole row segment */
                nrm = FN(nrm2)(kcol, row, lda);
            }
            T xi1 = row[(kcol - 1) * lda];
            if (nrm == (T)0) tk = (T)0;
            else {
                T nu = COPYSIGN(nrm, xi1);
                xi1 += nu;
                row[(kcol - 1) * lda] = -nu;
                for (int64_t i = 0; i < kcol - 1; ++i) row[i * lda] /= xi1;
                tk = xi1 / nu;
            }
        }
        tau[k - 1] = tk;
        /* reflectorYlppa!(x, tau, A[1:krow-1, 1:kcol]) (src/rq.jl:327-351) */
        for (int64_t j = 0; j < krow - 1; ++j) {
            T Ajv = A_(j, kcol - 1);
            for (int64_t i = 0; i < kcol - 1; ++i) Ajv += A_(j, i) * row[i * lda];
            Ajv = tk * Ajv;
            A_(j, kcol - 1) -= Ajv;
            for (int64_t i = 0; i < kcol - 1; ++i) A_(j, i) -= Ajv * row[i * lda];
        }
    }
}

/* lmul!(Q::RQPackedQ,B) generic loop (src/rq.jl:138-167, k descending) and lmul!(Q',B)
 * (src/rq.jl:203-233, k ascending).  factors mA x nA, B is nA x nB. */
int FN(rq_lmul)(int64_t mA, int64_t nA, const T *A, int64_t lda, const T *tau,
                int64_t mB, int64_t nB, T *B, int64_t ldb, int adjoint) {
    if (nA != mB) return -1;
    int64_t mm = mA < nA ? mA : nA;
    for (int64_t kk = 0; kk < mm; ++kk) {
        int64_t k = adjoint ? 1 + kk : mm - kk;
        int64_t krow = mA - mm + k, kcol = nA - mm + k;
        T tk = tau[k - 1];
        for (int64_t j = 0; j < nB; ++j) {
            T vBj = B_(kcol - 1, j);
            for (int64_t i = 0; i < kcol - 1; ++i) vBj += A_(krow - 1, i) * B_(i, j);
            vBj = tk * vBj;
            B_(kcol - 1, j) -= vBj;
            for (int64_t i = 0; i < kcol - 1; ++i) B_(i, j) -= A_(krow - 1, i) * vBj;
        }
    }
    return 0;
}

/* rmul!(C, Q::RQPackedQ) (src/rq.jl:243-271, k ascending) and rmul!(C, Q') (src/rq.jl:292-321,
 * k descending).  C is mC x nA. */
int FN(rq_rmul)(int64_t mA, int64_t nA, const T *A, int64_t lda, const T *tau,
                int64_t mC, int64_t nC, T *B, int64_t ldb, int adjoint) {
    if (nC != nA) return -1;
    int64_t mm = mA < nA ? mA : nA;
    for (int64_t kk = 0; kk < mm; ++kk) {
        int64_t k = adjoint ? mm - kk : 1 + kk;
        int64_t krow = mA - mm + k, kcol = nA - mm + k;
        T tk = tau[k - 1];
        for (int64_t i = 0; i < mC; ++i) {
            T v = B_(i, kcol - 1);
            for (int64_t j = 0; j < kcol - 1; ++j) v += B_(i, j) * A_(krow - 1, j);
            v = v * tk;
            B_(i, kcol - 1) -= v;
            for (int64_t j = 0; j < kcol - 1; ++j) B_(i, j) -= v * A_(krow - 1, j);
        }
    }
    return 0;
}

/*
 * UL factorization = reverse-order partially pivoted Gaussian elimination, generic_ulfact!
 * (src/ul.jl:148-196).  ipiv is 1-based like the reference (BlasInt), returns info.
 */
int64_t FN(ulfact)(int64_t m, int64_t n, T *A, int64_t lda, int64_t *ipiv, int pivot) {
    int64_t minmn = m < n ? m : n, info = 0;
    for (int64_t k = minmn; k >= 1; --k) {
        int64_t kp = k;
        if (pivot) {
            T amax = (T)0;
            for (int64_t i = 1; i <= k; ++i) {
                T absi = FABS(A_(i - 1, k - 1));
                if (absi > amax) { kp = i; amax = absi; }
            }
        }
        ipiv[k - 1] = kp;
        if (A_(kp - 1, k - 1) != (T)0) {
            if (k != kp)
                for (int64_t i = 0; i < n; ++i) {
                    T tmp = A_(k - 1, i); A_(k - 1, i) = A_(kp - 1, i); A_(kp - 1, i) = tmp;
                }
            T Akkinv = (T)1 / A_(k - 1, k - 1);
            for (int64_t i = 0; i < k - 1; ++i) A_(i, k - 1) *= Akkinv;
        } else if (info == 0) {
            info = k;
        }
        for (int64_t j = 0; j < k - 1; ++j)
            for (int64_t i = 0; i < k - 1; ++i) A_(i, j) -= A_(i, k - 1) * A_(k - 1, j);
    }
    return info;
}

/* ldiv!(F::UL, B) (src/ul.jl:390-393): rows permuted by ipiv in REVERSE order
 * (_apply_ipiv_rows!, src/ul.jl:363-376), then UnitUpper solve, then Lower solve.  Square n. */
int FN(ul_ldiv)(int64_t n, const T *A, int64_t lda, const int64_t *ipiv, int64_t nB, T *B, int64_t ldb) {
    for (int64_t i = n; i >= 1; --i)
        if (ipiv[i - 1] != i)
            for (int64_t c = 0; c < nB; ++c) {
                T t = B_(i - 1, c); B_(i - 1, c) = B_(ipiv[i - 1] - 1, c); B_(ipiv[i - 1] - 1, c) = t;
            }
    for (int64_t c = 0; c < nB; ++c) {
        for (int64_t i = n - 1; i >= 0; --i) {                  /* unit upper: back substitution */
            T s = B_(i, c);
            for (int64_t j = i + 1; j < n; ++j) s -= A_(i, j) * B_(j, c);
            B_(i, c) = s;
        }
        for (int64_t i = 0; i < n; ++i) {                       /* lower: forward substitution */
            T s = B_(i, c);
            for (int64_t j = 0; j < i; ++j) s -= A_(i, j) * B_(j, c);
            if (A_(i, i) == (T)0) return (int)(i + 1);
            B_(i, c) = s / A_(i, i);
        }
    }
    return 0;
}

/* Batched helpers: the reference-equivalent workload `for i: ql!(view(A,:,:,i))`. */
void FN(geql2_batched)(int64_t m, int64_t n, T *A, int64_t lda, int64_t strideA, T *tau,
                       int64_t strideTau, int64_t batch, int convention) {
    for (int64_t b = 0; b < batch; ++b)
        FN(geql2)(m, n, A + b * strideA, lda, tau + b * strideTau, convention);
}

#undef A_
#undef B_
