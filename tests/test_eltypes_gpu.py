"""GPU parity of the other BlasFloat element types of the ql! hook (reference src/ql.jl:72: every T<:BlasFloat goes to
LAPACK.geqlf!; element-type sweep of test/test_ql.jl:35-56): Float32, ComplexF64, ComplexF32 -- through the C-ABI, against
the LAPACK routines the reference calls (?geqlf, ?unmql/?ormql) and the relations the reference's tests pin.

Tolerances (BASELINE.json): L and tau elementwise within 50 n eps ||A||, ||QL - A||/||A|| and ||Q'Q - I|| within 20 n eps,
F\\b within the reference's 5000 eps class (test/test_ql.jl:69) scaled by cond(A).
"""
import numpy as np
import pytest

from oracle import lapack
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.complex128, np.complex64]


@pytest.fixture(scope="module")
def q():
    return load_package()


@pytest.fixture(scope="module")
def h(q):
    hd = q.Handle(0)
    yield hd
    hd.close()


def _rand(rng, shape, dt):
    a = rng.standard_normal(shape)
    if np.issubdtype(dt, np.complexfloating):
        a = a + 1j * rng.standard_normal(shape)
    return np.asfortranarray(a.astype(dt))


def _eps(dt):
    return np.finfo(dt).eps


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(1, 1), (5, 5), (33, 33), (64, 64), (100, 100), (257, 257), (300, 200), (200, 300), (1000, 1000),
                                   (70, 1), (1, 70), (640, 96)])
def test_ql_factors_and_tau_vs_lapack(q, h, dt, shape):
    m, n = shape
    rng = np.random.default_rng(m * 1000 + n)
    A = _rand(rng, (m, n), dt)
    F = q.ql(A, handle=h)
    Fo, tauo = lapack.geqlf(A)
    k = min(m, n)
    tol = 50 * max(m, n) * _eps(dt) * max(np.linalg.norm(A), 1.0)
    assert F.factors.dtype == dt and F.tau.dtype == dt
    assert np.isfinite(F.factors).all()
    # L (and the reflectors): the whole packed array is pinned by LAPACK's conventions
    assert np.abs(F.factors - Fo).max() <= tol, np.abs(F.factors - Fo).max()
    assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * _eps(dt)
    # the diagonal of L is real for complex element types (zlarfg: beta real) except the last (top-left) entry of a square matrix
    if np.issubdtype(dt, np.complexfloating) and k > 1:
        d = np.array([F.factors[m - k + i, n - k + i] for i in range(1, k)])
        assert np.abs(d.imag).max() == 0.0


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(40, 40), (130, 130), (150, 90), (90, 150)])
def test_q_orthogonal_and_reconstruction(q, h, dt, shape):
    """test/test_ql.jl:58-66: q'q = I, q*l = a."""
    m, n = shape
    rng = np.random.default_rng(7 * m + n)
    A = _rand(rng, (m, n), dt)
    F = q.ql(A, handle=h)
    k = min(m, n)
    eps = _eps(dt)
    Qm = q.lmul_(F.Q, np.asfortranarray(np.eye(m, dtype=dt)))              # the square Q
    assert np.abs(Qm.conj().T @ Qm - np.eye(m)).max() <= 20 * m * eps
    Lfull = np.zeros((m, n), dtype=dt)
    Lfull[m - k:, :] = F.L
    R = Qm @ Lfull
    assert np.linalg.norm(R - A) / np.linalg.norm(A) <= 20 * max(m, n) * eps


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(50, 50), (129, 129), (140, 60), (60, 140)])
@pytest.mark.parametrize("nc", [1, 3, 40])
def test_lmul_rmul_vs_lapack_unmql(q, h, dt, shape, nc):
    """lmul!(Q,B), lmul!(Q',B), rmul!(C,Q), rmul!(C,Q') (src/ql.jl:321-435 incl. the conj placements) vs LAPACK ?unmql / ?ormql."""
    m, n = shape
    rng = np.random.default_rng(m + 13 * n + nc)
    A = _rand(rng, (m, n), dt)
    Fo, tauo = lapack.geqlf(A)
    k = min(m, n)
    V = np.asfortranarray(Fo[:, n - k:])
    Qo = q.QLPackedQ(Fo, tauo, handle=h)
    adj = "C" if np.issubdtype(dt, np.complexfloating) else "T"
    tol = 50 * max(m, n) * _eps(dt) * np.sqrt(m)
    B = _rand(rng, (m, nc), dt)
    C = _rand(rng, (nc, m), dt)
    got = q.lmul_(Qo, B.copy(order="F"))
    assert np.abs(got - lapack.ormql("L", "N", V, tauo, B)).max() <= tol * max(1.0, np.abs(B).max())
    got = q.lmul_(Qo.T, B.copy(order="F"))
    assert np.abs(got - lapack.ormql("L", adj, V, tauo, B)).max() <= tol * max(1.0, np.abs(B).max())
    got = q.rmul_(C.copy(order="F"), Qo)
    assert np.abs(got - lapack.ormql("R", "N", V, tauo, C)).max() <= tol * max(1.0, np.abs(C).max())
    got = q.rmul_(C.copy(order="F"), Qo.T)
    assert np.abs(got - lapack.ormql("R", adj, V, tauo, C)).max() <= tol * max(1.0, np.abs(C).max())


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n", [1, 7, 64, 100, 333])
@pytest.mark.parametrize("nrhs", [1, 2, 9])
def test_solve_square(q, h, dt, n, nrhs):
    r"""a*(qla\b) ≈ b atol=5000ε (test/test_ql.jl:69), scaled by the conditioning of the random matrix."""
    rng = np.random.default_rng(n * 31 + nrhs)
    A = _rand(rng, (n, n), dt)
    b = _rand(rng, (n, nrhs), dt)
    F = q.ql(A, handle=h)
    x = F.solve(b)
    assert x.dtype == dt
    cond = np.linalg.cond(A.astype(np.complex128 if np.issubdtype(dt, np.complexfloating) else np.float64))
    assert np.abs(A @ x - b).max() <= 5000 * _eps(dt) * max(cond, 1.0)
    xo = np.linalg.solve(A.astype(np.complex128), b.astype(np.complex128))
    assert np.abs(x - xo).max() <= 50 * n * _eps(dt) * cond * max(1.0, np.abs(xo).max())


@pytest.mark.parametrize("dt", DTYPES)
def test_device_pointers_and_lda(q, h, dt):
    """Device-resident column-major tensors with lda > m give the same bits as the packed host call (test/test_ql.jl:76 pins
    storage independence)."""
    import torch

    rng = np.random.default_rng(5)
    m, n, lda = 120, 90, 131
    A = _rand(rng, (m, n), dt)
    F = q.ql(A, handle=h)
    tdt = {np.float32: torch.float32, np.complex128: torch.complex128, np.complex64: torch.complex64}[dt]
    buf = torch.zeros((n, lda), dtype=tdt, device="cuda")
    Ad = buf.T[:m, :]
    Ad.copy_(torch.from_numpy(A))
    Fd = q.ql_(Ad, handle=h)
    torch.cuda.synchronize()
    assert np.array_equal(Fd.factors.cpu().numpy(), F.factors)
    assert np.array_equal(Fd.tau.cpu().numpy(), F.tau)
    assert float(buf.T[m:, :].abs().max()) == 0.0          # the padding rows are untouched


@pytest.mark.parametrize("dt", [np.complex128, np.float32])
def test_extreme_scales_follow_larfg(q, h, dt):
    """Scaled norms + the safmin loop (LAPACK ?larfg semantics): matrices scaled far away from 1."""
    rng = np.random.default_rng(11)
    n = 48
    A = _rand(rng, (n, n), dt)
    scales = [1e-170, 1e150] if dt == np.complex128 else [1e-30, 1e30]
    for s in scales:
        As = np.asfortranarray((A * dt(s)).astype(dt))
        F = q.ql(As, handle=h)
        Fo, tauo = lapack.geqlf(As)
        assert np.isfinite(F.factors).all()
        nrm = float(np.linalg.norm(As.astype(np.complex128) / s)) * s          # (the Float32 sum of squares would underflow)
        D = np.abs(F.factors.astype(np.complex128) - Fo.astype(np.complex128))
        assert np.tril(D).max() <= 50 * n * _eps(dt) * nrm                     # L scales with A
        assert np.triu(D, 1).max() <= 50 * n * _eps(dt)                        # the reflector tails do not
        assert np.abs(np.tril(F.factors)).max() > 0.0
        assert np.abs(F.tau - tauo).max() <= 50 * n * _eps(dt)


def test_zero_tail_columns_complex(q, h):
    """zlarfg: tau = 0 only when the tail is zero AND alpha is real; a complex alpha with a zero tail still gets a (unitary) H."""
    A = np.zeros((4, 4), dtype=np.complex128, order="F")
    A[3, 3] = 2.0
    A[2, 2] = 1.0 + 1.0j
    A[1, 1] = -3.0
    A[0, 0] = 1.0j
    F = q.ql(A, handle=h)
    Fo, tauo = lapack.geqlf(A)
    assert np.abs(F.factors - Fo).max() <= 1e-14
    assert np.abs(F.tau - tauo).max() <= 1e-14


def test_dimension_errors_complex(q, h):
    A = _rand(np.random.default_rng(1), (6, 6), np.complex128)
    F = q.ql(A, handle=h)
    with pytest.raises(q.DimensionMismatch):
        q.lmul_(F.Q, np.zeros((5, 2), dtype=np.complex128, order="F"))
    with pytest.raises(q.DimensionMismatch):
        F.solve(np.zeros((5, 1), dtype=np.complex128, order="F"))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(40, 40), (130, 70), (70, 130), (1, 9), (9, 1)])
def test_rq_factors_and_q_apply_other_types(q, h, dt, shape):
    """rq!(A) = RQ(LAPACK.gerqf!(A)...) (src/rq.jl:401) and lmul!/rmul! with RQPackedQ (src/rq.jl:130-291 -> LAPACK.ormrq!) for
    Float32 / ComplexF64 / ComplexF32: ?gerqf(A) = (?geqlf(A^H))^H, ?unmrq through the conjugate-transposed reflectors."""
    m, n = shape
    rng = np.random.default_rng(m * 7 + n)
    A = _rand(rng, (m, n), dt)
    F, tau = q.rq_(A.copy(order="F"), handle=h)
    Fo, tauo = lapack.gerqf(A)
    tol = 50 * max(m, n) * _eps(dt) * max(np.linalg.norm(A), 1.0)
    assert np.abs(F - Fo).max() <= tol
    assert np.abs(tau - tauo).max() <= 50 * max(m, n) * _eps(dt)
    k = min(m, n)
    V = np.asfortranarray(Fo[m - k:, :])
    Qo = q.RQPackedQ(Fo, tauo, handle=h)
    adj = "C" if np.issubdtype(dt, np.complexfloating) else "T"
    B = _rand(rng, (n, 3), dt)
    C = _rand(rng, (3, n), dt)
    t2 = 50 * max(m, n) * _eps(dt) * np.sqrt(n)
    assert np.abs(q.lmul_(Qo, B.copy(order="F")) - lapack.ormrq("L", "N", V, tauo, B)).max() <= t2 * max(1.0, np.abs(B).max())
    assert np.abs(q.lmul_(Qo.T, B.copy(order="F")) - lapack.ormrq("L", adj, V, tauo, B)).max() <= t2 * max(1.0, np.abs(B).max())
    assert np.abs(q.rmul_(C.copy(order="F"), Qo) - lapack.ormrq("R", "N", V, tauo, C)).max() <= t2 * max(1.0, np.abs(C).max())
    assert np.abs(q.rmul_(C.copy(order="F"), Qo.T) - lapack.ormrq("R", adj, V, tauo, C)).max() <= t2 * max(1.0, np.abs(C).max())


@pytest.mark.parametrize("shape", [(2100, 2100), (2100, 2600), (2600, 2100)])
def test_complex_ql_lookahead_vs_lapack(q, h, shape):
    """The ComplexF64 ql! with the look-ahead schedule (panel j+1 on the panel SM partition while panel j is applied; automatic
    from min(m,n) = 8192 on, forced here with lookahead = 1): same parity bar against LAPACK zgeqlf as the sequential schedule."""
    m, n = shape
    rng = np.random.default_rng(m + 3 * n)
    A = _rand(rng, (m, n), np.complex128)
    Fo, tauo = lapack.geqlf(A)
    tol = 50 * max(m, n) * _eps(np.complex128) * np.linalg.norm(A)
    for la in (1, 0):
        h.set_option("lookahead", la)
        try:
            F = q.ql(A, handle=h)
        finally:
            h.set_option("lookahead", 2)
        assert np.abs(F.factors - Fo).max() <= tol
        assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * _eps(np.complex128)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(30, 30), (45, 28), (28, 45)])
def test_q_apply_and_solve_vs_the_reference_loops(q, h, dt, shape):
    """lmul!/rmul! (Q and Q') and the square ldiv! against the NumPy restatement of the reference's OWN generic loops
    (oracle/loops.py: src/ql.jl:318-435 with the conj placements, :440-447,467), on the GPU-computed factors."""
    from oracle import loops

    m, n = shape
    rng = np.random.default_rng(m * 3 + n)
    A = _rand(rng, (m, n), dt)
    F = q.ql(A, handle=h)
    hi = np.complex128 if np.issubdtype(dt, np.complexfloating) else np.float64
    Fh, th = F.factors.astype(hi), F.tau.astype(hi)
    B = _rand(rng, (m, 4), dt)
    C = _rand(rng, (4, m), dt)
    tol = 50 * max(m, n) * _eps(dt) * np.sqrt(m) * max(1.0, np.abs(B).max(), np.abs(C).max())
    assert np.abs(q.lmul_(F.Q, B.copy(order="F")) - loops.lmul(Fh, th, B.astype(hi))).max() <= tol
    assert np.abs(q.lmul_(F.Q.T, B.copy(order="F")) - loops.lmul(Fh, th, B.astype(hi), adjoint=True)).max() <= tol
    assert np.abs(q.rmul_(C.copy(order="F"), F.Q) - loops.rmul(C.astype(hi), Fh, th)).max() <= tol
    assert np.abs(q.rmul_(C.copy(order="F"), F.Q.T) - loops.rmul(C.astype(hi), Fh, th, adjoint=True)).max() <= tol
    if m == n:
        x = F.solve(B)
        xo = loops.ldiv_square(Fh, th, B.astype(hi))
        cond = np.linalg.cond(A.astype(hi))
        assert np.abs(x - xo).max() <= 50 * n * _eps(dt) * cond * max(1.0, np.abs(xo).max())
