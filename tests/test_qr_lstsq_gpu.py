"""SURVEY.md 8(f) rows N2 and N3, through the C-ABI:

  N3  qrunblocked!(A) = QR(LAPACK.geqrf!(A, tau)...) (reference src/qr.jl:72) and QRPackedQ's lmul!/rmul!, computed by the QL
      kernels through the flip identity geqrf(A) = J geqlf(J A J) J (test/test_ql.jl:7-13).  Oracle: LAPACK ?geqrf / ?ormqr
      (the routines the reference calls); tolerances as for QL: factors and tau within 50 n eps ||A||, ||QR - A||/||A|| and
      ||Q'Q - I|| within 20 n eps.
  N2  tall least squares F \\ b for m > n (the reference's own branch is @test_broken, test/test_ql.jl:148): checked against
      numpy.linalg.lstsq and the normal equations A'(Ax - b) = 0.
"""
import numpy as np
import pytest

from oracle import lapack
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def q():
    return load_package()


@pytest.fixture(scope="module")
def h(q):
    hd = q.Handle(0)
    yield hd
    hd.close()


@pytest.mark.parametrize("shape", [(1, 1), (10, 10), (12, 10), (10, 12), (7, 1), (1, 7), (64, 64), (200, 150), (150, 200),
                                   (129, 129), (513, 300), (1000, 1000), (2100, 1030)])
def test_qr_matches_geqrf(q, h, shape):
    m, n = shape
    rng = np.random.default_rng(m * 1000 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)) / 2)          # the reference's scale, test/test_qr.jl:11
    F, tau = q.qr_(A.copy(order="F"), handle=h)
    Fo, tauo = lapack.geqrf(A)
    nrm = np.linalg.norm(A)
    k = min(m, n)
    assert np.abs(F - Fo).max() <= 50 * max(m, n) * EPS * max(nrm, 1.0)
    assert np.abs(tau - tauo).max() <= 50 * max(m, n) * EPS
    # A = Q R with R = triu(factors)[:k], Q from the packed reflectors
    R = np.triu(F[:k, :])
    QR = np.zeros((m, n), order="F")
    QR[:k, :] = R
    QR = q.lmul_(q.QRPackedQ(F, tau, handle=h), QR)
    assert np.linalg.norm(QR - A) <= 20 * max(m, n) * EPS * max(nrm, 1e-300)
    E = np.asfortranarray(np.eye(m))
    Qm = q.lmul_(q.QRPackedQ(F, tau, handle=h), E)
    assert np.linalg.norm(Qm.T @ Qm - np.eye(m)) <= 20 * max(m, n) * EPS * np.sqrt(m)


@pytest.mark.parametrize("shape,nc", [((40, 40), 3), ((60, 25), 7), ((25, 60), 2), ((300, 200), 16)])
def test_qr_packed_q_apply_vs_ormqr(q, h, shape, nc):
    m, n = shape
    rng = np.random.default_rng(m + n + nc)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    Fo, tauo = lapack.geqrf(A)
    Q = q.QRPackedQ(Fo, tauo, handle=h)
    B = np.asfortranarray(rng.standard_normal((m, nc)))
    C = np.asfortranarray(rng.standard_normal((nc, m)))
    tol = 50 * m * EPS * max(np.abs(B).max(), np.abs(C).max()) * np.sqrt(m)
    for adj, tr in ((False, "N"), (True, "T")):
        Qa = Q.T if adj else Q
        assert np.abs(q.lmul_(Qa, B.copy(order="F")) - lapack.ormqr("L", tr, Fo, tauo, B)).max() <= tol
        assert np.abs(q.rmul_(C.copy(order="F"), Qa) - lapack.ormqr("R", tr, Fo, tauo, C)).max() <= tol
    with pytest.raises(q.DimensionMismatch):
        q.lmul_(Q, np.zeros((m + 1, 2), order="F"))


def test_qr_device_pointers_and_large(q, h):
    import torch

    n = 3000
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T
    F = (A.T.clone()).T
    Fd, tau = q.qr_(F, handle=h)
    torch.cuda.synchronize()
    Fo, tauo = lapack.geqrf(np.asfortranarray(A.cpu().numpy()))
    nrm = float(torch.linalg.norm(A))
    assert np.abs(Fd.cpu().numpy() - Fo).max() <= 50 * n * EPS * nrm
    assert np.abs(tau.cpu().numpy() - tauo).max() <= 50 * n * EPS
    # Q'A = R on the device
    QtA = q.lmul_(q.QRPackedQ(Fd, tau, handle=h).T, (A.T.clone()).T)
    torch.cuda.synchronize()
    assert float(torch.linalg.norm(torch.tril(QtA, diagonal=-1))) <= 20 * n * EPS * nrm
    assert float(torch.linalg.norm(torch.triu(QtA) - torch.triu(Fd))) <= 20 * n * EPS * nrm


@pytest.mark.parametrize("shape,nrhs", [((12, 10), 1), ((40, 7), 3), ((300, 120), 5), ((1500, 400), 2), ((700, 699), 4)])
def test_tall_least_squares(q, h, shape, nrhs):
    m, n = shape
    rng = np.random.default_rng(m + 7 * n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    b = np.asfortranarray(rng.standard_normal((m, nrhs)))
    F = q.ql(A, handle=h)
    x = F.solve(b)
    assert x.shape == (n, nrhs)
    xo = np.linalg.lstsq(A, b, rcond=None)[0]
    cond = np.linalg.cond(A)
    assert np.abs(x - xo).max() <= 50 * m * EPS * cond * max(1.0, np.abs(xo).max())
    r = A @ x - b
    assert np.abs(A.T @ r).max() <= 50 * m * EPS * np.linalg.norm(A) * max(np.linalg.norm(A) * np.abs(x).max(), np.abs(b).max()) * np.sqrt(m)
    # in-place form (what the reference's unchanged `\\` works on, src/ql.jl:496-501): the solution is B[1:n, :] -- the rows
    # `_cut_B(X, 1:n)` returns -- and the trailing m-n rows carry the residual's coordinates
    B = b.copy(order="F")
    q.ldiv_(F, B)
    assert np.allclose(np.linalg.norm(B[n:], axis=0), np.linalg.norm(r, axis=0), rtol=1e-9, atol=1e-12)
    assert np.array_equal(B[:n], x)
    v = F.solve(b[:, 0].copy())
    assert v.shape == (n,) and np.array_equal(v, x[:, 0])


@pytest.mark.parametrize("shape,nrhs", [((5, 9), 1), ((1, 6), 2), ((30, 47), 3), ((100, 101), 2), ((300, 1200), 5), ((1100, 1500), 2)])
def test_wide_minimum_norm_solve(q, h, shape, nrhs):
    r"""SURVEY.md 8(f) N2, wide half: `ql(A) \ b` for m < n returns the minimum-norm solution (what README.md:40 of the reference
    advertises; its own branch, src/ql.jl:448-483, is @test_broken at test/test_ql.jl:148): against numpy's lstsq (pinv solution)."""
    import torch

    m, n = shape
    rng = np.random.default_rng(m + 11 * n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    b = np.asfortranarray(rng.standard_normal((m, nrhs)))
    F = q.ql(A, handle=h)
    x = F.solve(b)
    assert x.shape == (n, nrhs)
    xo = np.linalg.lstsq(A, b, rcond=None)[0]
    cond = np.linalg.cond(A)
    assert np.abs(x - xo).max() <= 50 * n * EPS * cond * max(1.0, np.abs(xo).max())
    assert np.abs(A @ x - b).max() <= 5000 * EPS * cond * max(1.0, np.abs(b).max())
    # minimum norm: x lies in the row space of A (orthogonal to its null space)
    Q0 = np.linalg.qr(A.T, mode="complete")[0][:, m:]
    assert np.abs(Q0.T @ x).max() <= 50 * n * EPS * cond * max(1.0, np.abs(xo).max())
    v = F.solve(b[:, 0].copy())
    assert v.shape == (n,) and np.array_equal(v, x[:, 0])
    # device pointers, in-place form on the n-row buffer of src/ql.jl:496-501
    Fd = q.QL(torch.from_numpy(F.factors.T.copy()).cuda().T, torch.from_numpy(F.tau).cuda(), handle=h)
    X = torch.zeros((nrhs, n), dtype=torch.float64, device="cuda").T
    X[:m, :] = torch.from_numpy(b).cuda()
    q.ldiv_(Fd, X)
    torch.cuda.synchronize()
    assert np.array_equal(X.cpu().numpy(), x)
    with pytest.raises(q.DimensionMismatch):
        F.solve(np.ones(m + 1))


@pytest.mark.parametrize("shape,nrhs", [((10, 10), 1), ((64, 64), 3), ((300, 300), 4), ((40, 7), 2), ((1500, 400), 3)])
def test_qr_solve_square_and_tall(q, h, shape, nrhs):
    """`qrunblocked(A) \\ b` (src/qr.jl:193-204): exact solve / least squares through the flipped QL solve."""
    m, n = shape
    rng = np.random.default_rng(3 * m + n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    b = np.asfortranarray(rng.standard_normal((m, nrhs)))
    Fo, tauo = lapack.geqrf(A)
    x = q.qr_solve(Fo, tauo, b, handle=h)
    assert x.shape == (n, nrhs)
    xo = np.linalg.lstsq(A, b, rcond=None)[0]
    cond = np.linalg.cond(A)
    assert np.abs(x - xo).max() <= 50 * m * EPS * cond * max(1.0, np.abs(xo).max())
    if m == n:
        assert np.abs(A @ x - b).max() <= 5000 * EPS * cond
    # with our own QR factors too
    F, tau = q.qr_(A.copy(order="F"), handle=h)
    x2 = q.qr_solve(F, tau, b, handle=h)
    assert np.abs(x2 - xo).max() <= 50 * m * EPS * cond * max(1.0, np.abs(xo).max())


def test_qr_solve_singular_reports_r_index(q, h):
    A = np.asfortranarray(np.random.default_rng(1).standard_normal((20, 20)))
    Fo, tauo = lapack.geqrf(A)
    Fo[7, 7] = 0.0
    with pytest.raises(q.SingularException) as e:
        q.qr_solve(Fo, tauo, np.ones((20, 1), order="F"), handle=h)
    assert e.value.info == 8
