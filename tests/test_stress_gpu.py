"""Stateful stress of one handle: factorizations of changing sizes interleaved with lmul!/ldiv! on older and newer factors
(workspace regrowth, CUDA-graph cache, T-factor cache and its invalidation), every result checked against the oracle."""
import numpy as np
import pytest

from oracle import cport, lapack
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def test_interleaved_calls_on_one_handle():
    import torch

    q = load_package()
    h = q.Handle(0)
    rng = np.random.default_rng(2024)
    kept = []                                   # (A, F) pairs whose factors are reused later
    try:
        for it in range(40):
            m, n = int(rng.integers(1, 500)), int(rng.integers(1, 500))
            if it % 5 == 0:
                m = n = int(rng.integers(100, 700))
            A = np.asfortranarray(rng.standard_normal((m, n)))
            if it % 3 == 0:                     # device-pointer path, called twice with the same buffers (graph replay)
                Ad = torch.from_numpy(np.ascontiguousarray(A.T)).cuda().T
                Fd = q.ql_((Ad.T.clone()).T, handle=h)
                buf = Fd.factors
                buf.T.copy_(Ad.T)
                Fd = q.ql_(buf, handle=h)       # second sight: captured
                buf.T.copy_(Ad.T)
                Fd = q.ql_(buf, handle=h)       # third: replayed
                torch.cuda.synchronize()
                F = q.QL(np.asfortranarray(Fd.factors.cpu().numpy()), Fd.tau.cpu().numpy(), h)
                if m >= 2:
                    B = torch.from_numpy(np.ascontiguousarray(rng.standard_normal((2, m)))).cuda().T
                    got = q.lmul_(Fd.Q.T, (B.T.clone()).T)         # T cache hit on the device pointers
                    torch.cuda.synchronize()
                    ref = cport.lmul(F.factors, F.tau, np.asfortranarray(B.cpu().numpy()), adjoint=True)
                    assert np.abs(got.cpu().numpy() - ref).max() <= 50 * max(m, n) * EPS * max(1.0, np.abs(ref).max()) * np.sqrt(m)
            else:
                F = q.ql(A, handle=h)
            Fo, tauo = lapack.geqlf(A)
            tol = 50 * max(m, n) * EPS * max(np.linalg.norm(A), 1e-300)
            assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * EPS, (it, m, n)
            assert np.abs(np.asarray(F.factors) - Fo).max() <= tol, (it, m, n)
            kept.append((A, F))
            # reuse some OLDER factors: the T cache belongs to the newest factorization and must not be used for these
            A2, F2 = kept[int(rng.integers(0, len(kept)))]
            m2 = A2.shape[0]
            B = np.asfortranarray(rng.standard_normal((m2, 3)))
            got = q.lmul_(F2.Q if it % 2 else F2.Q.T, B.copy(order="F"))
            ref = cport.lmul(np.asfortranarray(F2.factors), F2.tau, B, adjoint=(it % 2 == 0))
            assert np.abs(got - ref).max() <= 50 * max(A2.shape) * EPS * max(1.0, np.abs(ref).max()) * np.sqrt(m2), (it, A2.shape)
            if A2.shape[0] == A2.shape[1] and A2.shape[0] > 1:
                x = F2.solve(B)
                assert np.abs(A2 @ x - B).max() <= 5000 * EPS * np.linalg.cond(A2) * max(1.0, np.abs(x).max())
    finally:
        h.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_larger_shapes_host_and_device(seed):
    """Random tall / wide / square shapes in the streaming range (>= 4M entries go through the chunked upload) and below,
    host and device entry points, against LAPACK."""
    import torch

    q = load_package()
    h = q.Handle(0)
    rng = np.random.default_rng(seed)
    try:
        for _ in range(3):
            m, n = int(rng.integers(900, 3600)), int(rng.integers(900, 3600))
            A = np.asfortranarray(rng.standard_normal((m, n)))
            lapack.set_num_threads(16)
            Fo, tauo = lapack.geqlf(A)
            tol = 50 * max(m, n) * EPS * np.linalg.norm(A)
            F = q.ql(A, handle=h)
            assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * EPS, (m, n)
            assert np.abs(F.factors - Fo).max() <= tol, (m, n)
            Ad = torch.from_numpy(np.ascontiguousarray(A.T)).cuda().T
            Fd = q.ql_(Ad, handle=h)
            torch.cuda.synchronize()
            assert np.abs(Fd.tau.cpu().numpy() - tauo).max() <= 50 * max(m, n) * EPS, (m, n)
            assert np.abs(Fd.factors.cpu().numpy() - Fo).max() <= tol, (m, n)
    finally:
        h.close()
