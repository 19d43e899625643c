"""Banded QL (reference ext/MatrixFactorizationsBandedMatricesExt.jl:14-36).

CPU part: the NumPy restatement of banded_ql! (oracle/banded.py) against dense LAPACK dgeqlf -- the relation the reference's own
tests assert (test/test_banded.jl:8-9,19-20,29-30,39-40,48-49,58-59,68-69,74-75).  GPU part: qlb200_dgbqlf through the C-ABI against
the same oracle and against dense LAPACK."""
import numpy as np
import pytest

from oracle import banded, lapack

SHAPES = [(10, 10, 3, 2), (14, 10, 3, 2), (10, 14, 3, 2), (10, 14, 3, 6), (100, 100, 3, 4), (102, 100, 3, 4), (100, 102, 3, 4),
          (50, 50, 1, 1), (64, 64, 0, 3), (40, 40, 5, 0), (1, 1, 0, 0), (300, 300, 17, 9)]


def _brand(rng, m, n, l, u):
    A = np.asfortranarray(rng.standard_normal((m, n)))
    i, j = np.indices((m, n))
    A[(i - j > l) | (j - i > u)] = 0.0
    return A


@pytest.mark.parametrize("m,n,l,u", SHAPES)
def test_oracle_banded_ql_matches_dense_lapack(m, n, l, u):
    rng = np.random.default_rng(m * 1000 + n * 10 + l + u)
    A = _brand(rng, m, n, l, u)
    lw, uw = banded.widened_bandwidths(m, n, l, u)
    D = banded.to_band(A, lw, uw)
    tau = banded.banded_ql(D, m, n, lw, uw)
    F = banded.from_band(D, m, n, lw, uw)
    tol = 50 * max(m, n) * np.finfo(float).eps * max(np.linalg.norm(A), 1.0)
    if u == 0:
        # no entries above the pivots: LinearAlgebra.reflector! on a length-1 vector flips the sign (tau = 2, beta = -alpha),
        # LAPACK dlarfg leaves it alone (tau = 0) -- the banded code follows the former
        k = min(m, n)
        assert np.all(tau[1 if m == n else 0:] == 2.0)
        assert np.abs(np.abs(F) - np.abs(A)).max() <= tol
        return
    Fo, tauo = lapack.geqlf(A)
    i, j = np.indices((m, n))
    inband = (i - j <= lw) & (j - i <= uw)
    assert np.abs(F - Fo)[inband].max() <= tol
    assert np.abs(Fo[~inband]).max(initial=0.0) <= tol          # the dense factors do not leave the widened band
    assert np.abs(tau - tauo).max() <= 50 * max(m, n) * np.finfo(float).eps


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,l,u", SHAPES + [(3000, 3000, 2, 2), (2000, 1900, 40, 30)])
def test_gpu_banded_ql_vs_oracle_and_dense_lapack(m, n, l, u):
    from __graft_entry__ import load_package

    q = load_package()
    h = q.default_handle(0)
    rng = np.random.default_rng(m + 7 * n + l + u)
    A = _brand(rng, m, n, l, u)
    lw, uw = banded.widened_bandwidths(m, n, l, u)
    D = banded.to_band(A, lw, uw)
    Dg, taug = q.banded_ql_(D.copy(order="F"), m, n, lw, uw, handle=h)
    tol = 50 * max(m, n) * np.finfo(float).eps * max(np.linalg.norm(A), 1.0)
    if m * n <= 20000:
        Do = D.copy(order="F")
        tauo = banded.banded_ql(Do, m, n, lw, uw)
        assert np.abs(Dg - Do).max() <= tol and np.abs(taug - tauo).max() <= 50 * max(m, n) * np.finfo(float).eps
    if u >= 1:
        Fo, tauL = lapack.geqlf(A)
        F = banded.from_band(Dg, m, n, lw, uw)
        i, j = np.indices((m, n))
        inband = (i - j <= lw) & (j - i <= uw)
        assert np.abs(F - Fo)[inband].max() <= tol
        assert np.abs(taug - tauL).max() <= 50 * max(m, n) * np.finfo(float).eps
    # device pointers
    import torch

    Dd = torch.from_numpy(D.T.copy()).cuda().T
    Dd2, taud = q.banded_ql_(Dd, m, n, lw, uw, handle=h)
    torch.cuda.synchronize()
    assert np.array_equal(Dd2.cpu().numpy(), Dg) and np.array_equal(taud.cpu().numpy(), taug)
