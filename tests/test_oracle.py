"""CPU tests (no GPU): pin the oracle (oracle/ql_oracle_impl.h C restatement) against
 (1) the LAPACK routine the reference calls (src/ql.jl:72, src/rq.jl:401) from scipy-openblas,
 (2) the reference's own relational known-answer tests (test/test_ql.jl cited per test),
 (3) the committed golden fixtures tests/golden/ql_golden.npz.
"""
import os

import numpy as np
import pytest

from oracle import cport, lapack

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ql_golden.npz"))
EPS = np.finfo(np.float64).eps


def _rng(seed=0):
    return np.random.default_rng(seed)


def tol(n, dt=np.float64, scale=1.0):
    return 50 * n * np.finfo(dt).eps * scale


@pytest.mark.parametrize("shape", [(10, 10), (10, 12), (12, 10), (3, 5), (64, 64), (100, 103), (200, 150), (1, 1), (5, 1), (1, 5)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_geql2_matches_lapack_geqlf(shape, dt):
    """test/test_ql.jl:5-33: LAPACK path == generic loop (factors and tau)."""
    A = (_rng(1).standard_normal(shape) / 2).astype(dt)
    F, tau = cport.geql2(A, convention=0)
    F2, tau2 = lapack.geqlf(A)
    t = tol(max(shape), dt, np.linalg.norm(A))
    assert np.abs(F - F2).max() <= t
    assert np.abs(tau - tau2).max() <= t
    # Julia generic path agrees on full-rank input (same sign convention)
    F3, tau3 = cport.geql2(A, convention=1)
    assert np.abs(F - F3).max() <= t and np.abs(tau - tau3).max() <= t


@pytest.mark.parametrize("shape", [(10, 10), (10, 12), (12, 10)])
def test_flip_qr_identity(shape):
    """test/test_ql.jl:7,12-13: R[end:-1:1,end:-1:1] == L, Q~[end:-1:1,end:-1:1] == Matrix(Q)."""
    A = _rng(2).standard_normal(shape)
    m, n = shape
    F, tau = cport.geql2(A)
    L = cport.getL(F)
    Fr, taur = lapack.geqrf(np.ascontiguousarray(A[::-1, ::-1]))
    R = np.triu(Fr)[: min(m, n), :]
    assert np.allclose(R[::-1, ::-1], L, atol=1e-12)
    assert np.allclose(taur[::-1], tau, atol=1e-12)
    Q = cport.square_q(F, tau)
    if m >= n:
        # Q L-block reconstructs A: A = Q[:, m-n:] @ L
        assert np.allclose(Q[:, m - n:] @ L, A, atol=1e-12)
    assert np.allclose(Q.T @ Q, np.eye(m), atol=1e-13)


def test_issue_7304_sign():
    """test/test_ql.jl:120-124: A=[-s -s; -s s] => Q == -A exactly-ish (norm(A+Q) < eps)."""
    A = GOLD["i7304_A"]
    F, tau = cport.geql2(A)
    Q = cport.matrix_q(F, tau)
    assert np.linalg.norm(A + Q) < 4 * EPS
    assert np.allclose(F, GOLD["i7304_F"], atol=4 * EPS) and np.allclose(tau, GOLD["i7304_tau"], atol=4 * EPS)


def test_zero_tail_conventions():
    """identity -> LAPACK tau = 0 (no flip); Julia generic tau = [0,2,2,2]."""
    F, tau = cport.geql2(np.eye(4), convention=0)
    assert np.all(tau == 0) and np.array_equal(F, np.eye(4))
    F2, tau2 = lapack.geqlf(np.eye(4))
    assert np.array_equal(tau2, tau) and np.array_equal(F2, F)
    _, tauj = cport.geql2(np.eye(4), convention=1)
    assert np.array_equal(tauj, [0, 2, 2, 2])


def test_wide_ql_reconstruction():
    """test/test_ql.jl:151-155 (3x5): Q*L == A."""
    A = GOLD["wide3x5_A"]
    F, tau = cport.geql2(A)
    L = cport.getL(F)
    assert np.allclose(cport.lmul(F, tau, L), A, atol=1e-13)


@pytest.mark.parametrize("shape", [(100, 100), (100, 103)])
def test_lmul_rmul_vs_matrix_q(shape):
    """test/test_ql.jl:157-181."""
    r = _rng(3)
    A = r.standard_normal(shape)
    F, tau = lapack.geqlf(A)
    Qs = cport.square_q(F, tau)
    x, b, c = r.standard_normal(100), r.standard_normal((100, 2)), r.standard_normal((2, 100))
    assert np.allclose(cport.lmul(F, tau, x), Qs @ x, atol=1e-12)
    assert np.allclose(cport.lmul(F, tau, b), Qs @ b, atol=1e-12)
    assert np.allclose(cport.lmul(F, tau, b, adjoint=True), Qs.T @ b, atol=1e-12)
    assert np.allclose(cport.rmul(c, F, tau), c @ Qs, atol=1e-12)
    assert np.allclose(cport.rmul(c, F, tau, adjoint=True), c @ Qs.T, atol=1e-12)
    # LAPACK's blocked ormql agrees with the scalar loops
    assert np.allclose(lapack.ormql("L", "T", F[:, -min(shape):] if shape[1] > shape[0] else F, tau, b),
                       Qs.T @ b, atol=1e-12)
    with pytest.raises(cport.DimensionMismatch):
        cport.lmul(F, tau, np.zeros((101, 2)))


def test_square_solve_tolerance():
    """test/test_ql.jl:69: a*(qla\\b) ~ b, atol = 5000 eps."""
    r = _rng(4)
    A = r.standard_normal((10, 10)) / 2
    b = r.standard_normal((10, 2))
    F, tau = cport.geql2(A)
    X = cport.ldiv(F, tau, b)
    assert np.abs(A @ X - b).max() < 5000 * EPS
    assert np.allclose(X, np.linalg.solve(A, b), atol=1e-10)


def test_lda_independence_bitwise():
    """test/test_ql.jl:76 uses ==: results must not depend on the leading dimension."""
    r = _rng(5)
    big = np.asfortranarray(r.standard_normal((12, 12)))
    view = big[:9, :9]                       # lda = 12
    F1, t1 = cport.geql2(view)               # _f copies to contiguous
    F2, t2 = cport.geql2(np.array(view, order="F"))
    assert np.array_equal(F1, F2) and np.array_equal(t1, t2)


@pytest.mark.parametrize("shape", [(10, 10), (8, 13), (40, 12)])
def test_rq_matches_gerqf_and_transposed_ql(shape):
    """test/test_rq.jl:29-67 (direct gerqf comparison) + SURVEY A.4: gerqf(A) == geqlf(A^T)^T."""
    A = _rng(6).standard_normal(shape)
    F, tau = cport.gerq2(A)
    F2, tau2 = lapack.gerqf(A)
    assert np.allclose(F, F2, atol=1e-12) and np.allclose(tau, tau2, atol=1e-12)
    F3, tau3 = lapack.geqlf(np.asfortranarray(A.T))
    assert np.allclose(F3.T, F2, atol=1e-12) and np.allclose(tau3, tau2, atol=1e-12)


@pytest.mark.parametrize("n", [1, 7, 33])
def test_ul_is_flipped_getrf(n):
    """ul(A) == flip(getrf(flip(A))), ipiv_UL[k] = n+1-ipiv_LU[n+1-k]; U*L == A[p,:]."""
    A = _rng(7).standard_normal((n, n))
    F, ipiv, info = cport.ulfact(A)
    LU, piv, info2 = lapack.getrf(np.ascontiguousarray(A[::-1, ::-1]))
    assert info == 0 and info2 == 0
    assert np.allclose(F, LU[::-1, ::-1], atol=1e-12)
    assert np.array_equal(ipiv, (n + 1 - piv)[::-1])
    b = _rng(8).standard_normal((n, 2))
    assert np.allclose(cport.ul_ldiv(F, ipiv, b), np.linalg.solve(A, b), atol=1e-9)


def test_ul_singular_info():
    """test/test_ul.jl:192-216: singular => info > 0."""
    A = np.zeros((3, 3))
    _, _, info = cport.ulfact(A)
    assert info > 0


def test_golden_fixtures():
    """Oracle vs committed goldens (LAPACK-generated factors/tau; loop-generated apply/solve)."""
    for name in ["sq10", "wide10x12", "tall12x10", "wide3x5", "sq100", "wide100x103", "sq200", "tall300x130", "sq64_f32"]:
        A, Fg, tg = GOLD[f"{name}_A"], GOLD[f"{name}_F"], GOLD[f"{name}_tau"]
        F, tau = cport.geql2(A)
        t = tol(max(A.shape), A.dtype, np.linalg.norm(A))
        assert np.abs(F - Fg).max() <= t, name
        assert np.abs(tau - tg).max() <= t, name
        if f"{name}_B" in GOLD:
            B = GOLD[f"{name}_B"]
            assert np.abs(cport.lmul(Fg, tg, B, adjoint=True) - GOLD[f"{name}_QtB"]).max() == 0, name
            assert np.abs(cport.ldiv(Fg, tg, B) - GOLD[f"{name}_X"]).max() == 0, name
    for name in ["rq_sq10", "rq_wide8x13", "rq_tall40x12"]:
        F, tau = cport.gerq2(GOLD[f"{name}_A"])
        assert np.allclose(F, GOLD[f"{name}_F"], atol=1e-12) and np.allclose(tau, GOLD[f"{name}_tau"], atol=1e-12)
    for name in ["ul10", "ul33"]:
        F, ipiv, _ = cport.ulfact(GOLD[f"{name}_A"])
        assert np.array_equal(F, GOLD[f"{name}_F"]) and np.array_equal(ipiv, GOLD[f"{name}_ipiv"])


@pytest.mark.parametrize("name,dt", [("b32_f64", np.float64), ("b16_f32", np.float32), ("b8_f64", np.float64)])
def test_batched_oracle_vs_golden(name, dt):
    A, Fg, tg = GOLD[f"{name}_A"], GOLD[f"{name}_F"], GOLD[f"{name}_tau"]
    F, tau = cport.geql2_batched(A)
    n = A.shape[1]
    t = tol(n, dt, np.linalg.norm(A, axis=(1, 2)).max())
    assert np.abs(F - Fg).max() <= t and np.abs(tau - tg).max() <= t


def test_larfg_edge_cases():
    """dlarfg conventions: zero tail -> tau=0, beta=alpha; tiny values rescale loop."""
    beta, v, tau = cport.larfg(3.0, np.zeros(4))
    assert tau == 0 and beta == 3.0
    beta, v, tau = cport.larfg(3.0, np.array([4.0]))
    assert np.isclose(beta, -5.0) and np.isclose(tau, 1.6) and np.isclose(v[0], 0.5)
    beta, v, tau = cport.larfg(1e-310, np.array([1e-310, 1e-310]))
    assert np.isfinite(tau) and 1.0 <= tau <= 2.0 and np.isclose(beta, -np.sqrt(3) * 1e-310, rtol=1e-6)
