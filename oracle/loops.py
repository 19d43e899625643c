"""NumPy restatement of the reference's generic Q-apply and solve loops for ANY element type (real or complex) -- the loops that
MatrixFactorizations.jl runs for every QLPackedQ, LAPACK-factored or not.

TEST INFRASTRUCTURE ONLY.  Line by line after /root/reference/src/ql.jl:
  lmul           :318-345   materialize!(::Lmul{<:QLPackedQLayout})      B <- Q B     (k ascending, conj(A[i,k]), tau[k])
  lmul_adjoint   :348-377   materialize!(::Lmul{<:AdjQLPackedQLayout})   B <- Q' B    (k descending, conj(tau[k]))
  rmul           :380-406   materialize!(::Rmul{<:Any,<:QLPackedQLayout})     A <- A Q   (k descending)
  rmul_adjoint   :409-435   materialize!(::Rmul{<:Any,<:AdjQLPackedQLayout})  A <- A Q'  (k ascending, conj(tau[k]))
  ldiv_square    :440-447,467   lmul!(Q', B) then the lower-triangular solve with L = tril(factors)
The inner i/j loops are vectorised (same arithmetic per entry up to summation order).  The real-typed C port of the same loops is
oracle/ql_oracle_impl.h; this file adds the complex element types of SURVEY.md 8(f) N1 (the conj placements of :336,366,368,400,426,429).
"""
import numpy as np


def _k_range(mA, nA):
    return range(max(nA - mA + 1, 1), nA + 1)          # 1-based, like the reference


def lmul(factors, tau, B, adjoint=False):
    A = np.asarray(factors)
    B = np.array(B, dtype=np.result_type(A, B), order="F", copy=True)
    vec = B.ndim == 1
    if vec:
        B = B.reshape(-1, 1)
    mA, nA = A.shape
    if B.shape[0] != mA:
        raise ValueError("DimensionMismatch")
    ks = list(_k_range(mA, nA))
    if adjoint:
        ks = ks[::-1]
    for k in ks:
        mu = mA + k - nA
        v = A[: mu - 1, k - 1]
        t = tau[k - nA + min(mA, nA) - 1]
        vB = B[mu - 1, :] + np.conj(v) @ B[: mu - 1, :]
        vB = (np.conj(t) if adjoint else t) * vB
        B[mu - 1, :] -= vB
        B[: mu - 1, :] -= np.outer(v, vB)
    return B[:, 0] if vec else B


def rmul(Amat, factors, tau, adjoint=False):
    Q = np.asarray(factors)
    A = np.array(Amat, dtype=np.result_type(Q, Amat), order="F", copy=True)
    mQ, nQ = Q.shape
    if A.shape[1] != mQ:
        raise ValueError("DimensionMismatch")
    ks = list(_k_range(mQ, nQ))
    if not adjoint:
        ks = ks[::-1]
    for k in ks:
        mu = mQ + k - nQ
        v = Q[: mu - 1, k - 1]
        t = tau[k - nQ + min(mQ, nQ) - 1]
        vA = A[:, mu - 1] + A[:, : mu - 1] @ v
        vA = vA * (np.conj(t) if adjoint else t)
        A[:, mu - 1] -= vA
        A[:, : mu - 1] -= np.outer(vA, np.conj(v))
    return A


def ldiv_square(factors, tau, B):
    A = np.asarray(factors)
    n = A.shape[0]
    assert A.shape == (n, n)
    X = lmul(A, tau, B, adjoint=True)
    L = np.tril(A)
    import scipy.linalg

    return scipy.linalg.solve_triangular(L, X, lower=True)
