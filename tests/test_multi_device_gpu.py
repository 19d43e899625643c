"""Second-device coverage (skipped on a one-GPU box): every entry point run through a handle on cuda:1 must give
the same bits as the handle on cuda:0 -- the kernels are deterministic (fixed-order split-K reduce, no atomics), and
per-kernel attributes (dynamic shared memory, non-portable cluster size) are set per handle, i.e. per device.
Host-pointer C-ABI calls, so no torch device state is involved."""
import numpy as np
import pytest

from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu


def _two_devices():
    try:
        import torch

        return torch.cuda.is_available() and torch.cuda.device_count() >= 2
    except Exception:
        return False


needs2 = pytest.mark.skipif(not _two_devices(), reason="needs two GPUs")


@pytest.fixture(scope="module")
def q():
    return load_package()


@pytest.fixture(scope="module")
def hs(q):
    h0, h1 = q.Handle(0), q.Handle(1)
    yield h0, h1
    h0.close()
    h1.close()


@needs2
def test_ql_lmul_ldiv_same_bits_on_second_device(q, hs):
    rng = np.random.default_rng(7)
    n = 1500
    A = np.asfortranarray(rng.standard_normal((n, n)))
    B = np.asfortranarray(rng.standard_normal((n, 5)))
    out = []
    for h in hs:
        F = q.ql_(A.copy(order="F"), handle=h)
        QtB = q.lmul_(F.Q.T, B.copy(order="F"))
        X = q.ldiv_(F, B.copy(order="F"))
        out.append((F.factors, F.tau, QtB, X))
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    assert np.linalg.norm(A @ out[1][3] - B) <= 1e-9 * np.linalg.norm(B)


@needs2
def test_ul_and_rq_same_bits_on_second_device(q, hs):
    rng = np.random.default_rng(8)
    n = 700
    A = np.asfortranarray(rng.standard_normal((n, n)))
    b = np.asfortranarray(rng.standard_normal((n, 3)))
    out = []
    for h in hs:
        F = q.ul_(A.copy(order="F"), handle=h)
        x = F.solve(b.copy(order="F"))
        R, tau = q.rq_(np.asfortranarray(A[:300, :]).copy(order="F"), handle=h)
        out.append((F.factors, F.ipiv, x, R, tau))
    for a, b_ in zip(*out):
        assert np.array_equal(a, b_)
    assert np.linalg.norm(A @ out[1][2] - b) <= 1e-9 * np.linalg.norm(b)


@needs2
@pytest.mark.parametrize("n,dt", [(32, np.float64), (16, np.float32), (8, np.float64), (24, np.float64)])
def test_batched_and_sharded_same_bits(q, hs, n, dt):
    rng = np.random.default_rng(9)
    batch = 20001                     # odd: the two shards differ in size
    A = rng.standard_normal((batch, n, n)).astype(dt)
    B = rng.standard_normal((batch, 4, n)).astype(dt)
    F0, t0 = q.ql_batched_(A.copy(), handle=hs[0])
    F1, t1 = q.ql_batched_(A.copy(), handle=hs[1])
    Fm, tm = q.ql_batched_mg_(A.copy(), handles=list(hs))
    assert np.array_equal(F0, F1) and np.array_equal(t0, t1)
    assert np.array_equal(F0, Fm) and np.array_equal(t0, tm)
    W0 = q.lmul_batched_(F0, t0, B.copy(), adjoint=True, handle=hs[0])
    W1 = q.lmul_batched_(F0, t0, B.copy(), adjoint=True, handle=hs[1])
    X0 = q.ldiv_batched_(F0, t0, B.copy(), handle=hs[0])
    X1 = q.ldiv_batched_(F0, t0, B.copy(), handle=hs[1])
    assert np.array_equal(np.asarray(W0), np.asarray(W1))
    assert np.array_equal(np.asarray(X0), np.asarray(X1))
