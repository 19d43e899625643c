"""CPU restatement (NumPy loops) of the banded QL of the reference's BandedMatrices extension.

TEST INFRASTRUCTURE ONLY (never imported by the product).  Follows /root/reference/ext/MatrixFactorizationsBandedMatricesExt.jl:
  :14      ql(A::BandedMatrix) widens the lower bandwidth to max(l, l + u + m - n) before factoring,
  :18-36   banded_ql!(L): for k = n ... max(n-m+1+(T<:Real), 1): reflector! on the reversed band column, reflectorApply! on the
           columns j = max(1, mu-l) ... k-1.
Band storage = BandedMatrices.jl's `bandeddata`: D is (l+u+1) x n with A[i, j] = D[u + i - j, j] (0-based).
Pinned by tests/test_banded.py against dense LAPACK dgeqlf, the relation the reference's own tests assert
(test/test_banded.jl:8-9,19-20,29-30,39-40,48-49: `ql(A).factors ≈ ql!(Matrix(A)).factors`, same for tau).
"""
import numpy as np


def widened_bandwidths(m, n, l, u):
    """ext:14"""
    return max(l, l + u + m - n), u


def to_band(A, l, u):
    """Dense m x n -> band data ((l+u+1) x n, Fortran order); entries outside the band are dropped."""
    m, n = A.shape
    D = np.zeros((l + u + 1, n), dtype=A.dtype, order="F")
    for j in range(n):
        for i in range(max(0, j - u), min(m, j + l + 1)):
            D[u + i - j, j] = A[i, j]
    return D


def from_band(D, m, n, l, u):
    A = np.zeros((m, n), dtype=D.dtype, order="F")
    for j in range(n):
        for i in range(max(0, j - u), min(m, j + l + 1)):
            A[i, j] = D[u + i - j, j]
    return A


def reflector(x):
    """LinearAlgebra.reflector! (real case): x[0] is the pivot.  Returns tau; x[0] <- -nu, x[1:] <- x[1:] / (x[0] + nu)."""
    if x.size == 0:
        return 0.0
    xi1 = x[0]
    normu = np.linalg.norm(x)
    if normu == 0:
        return 0.0
    nu = np.copysign(normu, xi1)
    xi1 = xi1 + nu
    x[0] = -nu
    x[1:] /= xi1
    return xi1 / nu


def banded_ql(D, m, n, l, u):
    """ext:18-36 on the band data (in place).  Returns tau (min(m,n))."""
    tau = np.zeros(min(m, n), dtype=D.dtype)
    kmin = max(n - m + 1 + 1, 1)                   # T <: Real
    for k in range(n, kmin - 1, -1):               # 1-based like the reference
        mu = m + k - n
        rows = np.arange(u + 1 + mu - k, max(1, u - k + 2) - 1, -1) - 1       # data rows of column k, pivot first (0-based)
        if rows.size and rows[0] < 0:
            rows = rows[:0]
        x = D[rows, k - 1].copy()
        tk = reflector(x)
        D[rows, k - 1] = x
        tau[k - n + min(m, n) - 1] = tk
        N = x.size
        if N == 0:                                 # the pivot lies above the band (wide matrices): empty views, nothing to do
            continue
        for j in range(max(1, mu - l), k):
            rj = np.arange(u + 1 + mu - j, u + 2 + mu - j - N - 1, -1) - 1
            a = D[rj, j - 1]
            v = a[0] + x[1:] @ a[1:]
            v = tk * v
            a[0] -= v
            a[1:] -= x[1:] * v
            D[rj, j - 1] = a
    return tau
