"""GPU parity tests of the UL factorization (A13), RQ's packed Q (A12) and the multi-device batched entry
points through the C-ABI vs the oracle.

UL: the reference is generic_ulfact! (src/ul.jl:148-196) restated in oracle/ql_oracle_impl.h and pinned
against flipped LAPACK getrf in tests/test_oracle.py.  The pivot sequence must match EXACTLY (integer
work: bit-exact); factors within 50*n*eps*||A|| (the blocked GPU elimination applies updates in another
order than the reference's scalar loop).
RQ Q-apply: reference = LAPACK.ormrq! (src/rq.jl:130-137,183-202,234-242,272-291) and the C restatement
of the generic loops (src/rq.jl:138-167,203-233).
"""
import os

import numpy as np
import pytest

from oracle import cport, lapack
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ql_golden.npz"))
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def q():
    return load_package()


@pytest.fixture(scope="module")
def h(q):
    hd = q.Handle(0)
    yield hd
    hd.close()


@pytest.mark.parametrize("n", [1, 2, 7, 16, 17, 33, 100, 257, 600])
@pytest.mark.parametrize("pivot", [True, False])
def test_ul_matches_reference_loops(q, h, n, pivot):
    rng = np.random.default_rng(n * 3 + int(pivot))
    A = np.asfortranarray(rng.standard_normal((n, n)))
    if not pivot:
        A += n * np.eye(n)           # keep the unpivoted elimination well conditioned
    F = q.ul(A, pivot=pivot, handle=h)
    Fo, ipo, info = cport.ulfact(A, pivot)
    assert info == 0 and F.info == 0
    assert np.array_equal(np.asarray(F.ipiv), ipo), "pivot sequence differs from generic_ulfact!"
    growth = max(1.0, np.abs(Fo).max() / max(np.abs(A).max(), 1e-300))
    assert np.abs(F.factors - Fo).max() <= 50 * n * EPS * np.linalg.norm(A) * growth
    # A[p,:] = U*L (test/test_ul.jl:30-35)
    assert np.abs(A[F.p, :] - F.U @ F.L).max() <= 50 * n * EPS * np.linalg.norm(A) * growth


@pytest.mark.parametrize("name", ["ul10", "ul33"])
def test_ul_golden(q, h, name):
    A = GOLD[f"{name}_A"]
    F = q.ul(A, handle=h)
    assert np.array_equal(np.asarray(F.ipiv), GOLD[f"{name}_ipiv"])
    assert np.abs(F.factors - GOLD[f"{name}_F"]).max() <= 50 * A.shape[0] * EPS * np.linalg.norm(A)


@pytest.mark.parametrize("shape", [(40, 12), (12, 40), (100, 33)])
def test_ul_rectangular_follows_the_reference_loop(q, h, shape):
    """Rectangular inputs: only rows/columns 1..min(m,n) are eliminated, interchanges span all n columns
    (src/ul.jl:155-191; the reference's own thin/fat tests are commented out, test/test_ul.jl:174-181)."""
    rng = np.random.default_rng(shape[0])
    A = np.asfortranarray(rng.standard_normal(shape))
    F = q.ul(A, handle=h)
    Fo, ipo, info = cport.ulfact(A, True)
    assert np.array_equal(np.asarray(F.ipiv), ipo)
    assert np.abs(F.factors - Fo).max() <= 50 * max(shape) * EPS * np.linalg.norm(A) * max(1.0, np.abs(Fo).max())


def test_ul_ties_pick_the_first_maximum_and_singular_info(q, h):
    # ties: the reference's strict '>' scan keeps the FIRST row attaining the maximum (src/ul.jl:162)
    A = np.asfortranarray(np.array([[2.0, 1.0, 3.0], [1.0, 5.0, -3.0], [0.5, 2.0, 3.0]]))
    F = q.ul(A, handle=h)
    _, ipo, _ = cport.ulfact(A, True)
    assert np.array_equal(np.asarray(F.ipiv), ipo) and ipo[2] == 1
    # an exactly singular matrix: info = the first zero pivot the descending loop meets; check=True raises
    S = np.asfortranarray(np.array([[1.0, 2.0, 0.0], [3.0, 4.0, 0.0], [5.0, 6.0, 0.0]]))
    Fs = q.ul(S, check=False, handle=h)
    _, _, info = cport.ulfact(S, True)
    assert Fs.info == info == 3
    with pytest.raises(q.SingularException):
        q.ul(S, handle=h)


@pytest.mark.parametrize("n,nrhs", [(5, 1), (33, 3), (200, 8), (300, 17)])
def test_ul_solve(q, h, n, nrhs):
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, n)))
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    F = q.ul(A, handle=h)
    X = F.solve(B)
    Fo, ipo, _ = cport.ulfact(A, True)
    Xo = cport.ul_ldiv(Fo, ipo, B)
    cond = np.linalg.cond(A)
    assert np.abs(X - Xo).max() <= 50 * n * EPS * cond * max(1.0, np.abs(Xo).max())
    assert np.abs(A @ X - B).max() <= 5000 * EPS * cond          # test/test_ul.jl solve tolerance class
    # vector right-hand side
    x = F.solve(B[:, 0].copy())
    assert np.abs(x - X[:, 0]).max() == 0.0


def test_ul_device_pointers(q, h):
    import torch

    n = 384
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T
    A0 = A.clone()
    F = q.ul_(A, handle=h)
    torch.cuda.synchronize()
    Fo, ipo, _ = cport.ulfact(A0.cpu().numpy(), True)
    assert np.array_equal(F.ipiv.cpu().numpy(), ipo)
    assert np.abs(F.factors.cpu().numpy() - Fo).max() <= 50 * n * EPS * np.linalg.norm(Fo)


@pytest.mark.parametrize("shape", [(10, 10), (8, 13), (40, 12), (130, 300)])
@pytest.mark.parametrize("nrhs", [1, 5])
def test_rq_q_apply_vs_ormrq(q, h, shape, nrhs):
    m, n = shape
    rng = np.random.default_rng(m * 100 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F, tau = lapack.gerqf(A)
    k = min(m, n)
    Q = q.RQPackedQ(F, tau, handle=h)
    B = np.asfortranarray(rng.standard_normal((n, nrhs)))
    C = np.asfortranarray(rng.standard_normal((nrhs, n)))
    tol = 50 * n * EPS * max(np.abs(B).max(), np.abs(C).max()) * n
    Vk = np.asfortranarray(F[m - k:, :])        # the k reflector rows
    for adj, tr in [(False, "N"), (True, "T")]:
        got = q.lmul_(Q.T if adj else Q, B.copy(order="F"))
        ref = lapack.ormrq("L", tr, Vk, tau, B.copy(order="F"))
        assert np.abs(got - ref).max() <= tol, ("L", tr)
        assert np.abs(got - cport.rq_lmul(F, tau, B, adjoint=adj)).max() <= tol
        got = q.rmul_(C.copy(order="F"), Q.T if adj else Q)
        ref = lapack.ormrq("R", tr, Vk, tau, C.copy(order="F"))
        assert np.abs(got - ref).max() <= tol, ("R", tr)
        assert np.abs(got - cport.rq_rmul(C, F, tau, adjoint=adj)).max() <= tol
    with pytest.raises(q.DimensionMismatch):
        q.lmul_(Q, np.zeros((n + 1, 2), order="F"))


def test_rq_tall_skinny_config4_properties(q, h):
    """configs[3] family (rq of a tall-skinny matrix), reduced rows so that the oracle finishes in seconds:
    gerqf(A) with m >> n touches only the last n rows' reflectors; compare with LAPACK dgerqf."""
    m, n = 4096, 128
    rng = np.random.default_rng(4)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F, tau = q.rq_(A.copy(order="F"), handle=h)
    Fo, tauo = lapack.gerqf(A)
    tol = 50 * m * EPS * np.linalg.norm(A)
    assert np.abs(np.triu(F[m - n:, :]) - np.triu(Fo[m - n:, :])).max() <= tol
    assert np.abs(tau - tauo).max() <= 50 * m * EPS


def test_batched_mg_single_device_sharding(q, h):
    """The `_mg` entry point with two handles on the same device: shards [0,b/2) and [b/2,b) run from two
    host threads; results equal the single-handle call bit for bit (same kernels, same data)."""
    h2 = q.Handle(0)
    try:
        rng = np.random.default_rng(9)
        A = rng.standard_normal((1001, 32, 32))
        F1, t1 = q.ql_batched_(A.copy(), handle=h)
        F2, t2 = q.ql_batched_mg_(A.copy(), handles=[h, h2])
        assert np.array_equal(F1, F2) and np.array_equal(t1, t2)
    finally:
        h2.close()


@pytest.mark.parametrize("shape", [(2500, 2500), (3000, 2300), (2300, 3000)])
@pytest.mark.parametrize("pivot", [True, False])
def test_ul_lookahead_same_bits_and_flipped_getrf(q, h, shape, pivot):
    """From min(m,n) >= 2048 on ul! factors block b+1 on the panel SM partition while block b is applied (look-ahead, row
    interchanges outside the block deferred): same operations in the same order per entry, so the same bits as the
    sequential schedule (option lookahead = 0); pivots and factors vs LAPACK getrf of the flipped matrix (SURVEY.md A.4,
    test/test_ul.jl:30-35) for the square pivoted case."""
    import torch

    m, n = shape
    rng = np.random.default_rng(m + n + int(pivot))
    A = np.asfortranarray(rng.standard_normal((m, n)) + (0 if pivot else 4.0) * np.eye(m, n))
    Ad = torch.empty_strided((m, n), (1, m), dtype=torch.float64, device="cuda").copy_(torch.from_numpy(A))
    F1 = q.ul_(Ad.clone(memory_format=torch.preserve_format), pivot=pivot, check=False, handle=h)
    h.set_option("lookahead", 0)
    try:
        F0 = q.ul_(Ad.clone(memory_format=torch.preserve_format), pivot=pivot, check=False, handle=h)
    finally:
        h.set_option("lookahead", 2)
    torch.cuda.synchronize()
    assert torch.equal(F1.factors, F0.factors) and torch.equal(F1.ipiv, F0.ipiv) and F1.info == F0.info
    if pivot and m == n:
        LU, ipiv, info = lapack.getrf(A[::-1, ::-1])
        want = LU[::-1, ::-1]
        got = F1.factors.cpu().numpy()
        assert np.array_equal(F1.ipiv.cpu().numpy(), (n + 1 - ipiv)[::-1])
        assert np.abs(got - want).max() <= 50 * n * EPS * np.linalg.norm(A)
