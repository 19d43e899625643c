"""GPU tests at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot factor
these sizes in seconds, so a sample is compared with it and the rest is checked by invariants):

  configs[1]  ql! Float64 n = 16384: Q [0;L] = A and Q'Q = I on sampled columns (20*n*eps), tau range;
  configs[2]  262144 x (32x32) Float64 + lmul!(Q',B): sampled matrices vs the oracle (50*n*eps*||A||),
              lmul!(Q, lmul!(Q', B)) round trip and norm preservation for EVERY matrix;
  configs[3]  rq of 65536 x 1024 Float64 vs LAPACK dgerqf (the routine behind rq!, src/rq.jl:401);
  configs[4]  1048576 x (16x16) Float32 + ldiv! on 8 right-hand sides: sampled matrices vs the oracle, backward
              error ||A x - b|| <= c n eps ||A|| ||x|| for EVERY matrix.
All calls go through the C-ABI `_dev` entry points on torch device tensors.
"""
import numpy as np
import pytest

from oracle import cport, lapack
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


def test_config2_ql_n16384_properties(q, h):
    import torch

    n, ns = 16384, 64
    g = torch.Generator(device="cuda").manual_seed(1234321)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T          # column-major
    Fm = (A.T.clone()).T
    F = q.ql_(Fm, handle=h)
    torch.cuda.synchronize()
    tau = F.tau
    assert bool(torch.isfinite(F.factors).all()) and bool(torch.isfinite(tau).all())
    assert float(tau[1:].min()) >= 1.0 and float(tau.max()) <= 2.0 and float(tau[0]) == 0.0   # real ?larfg: tau in [1,2], tau_1 = 0
    cols = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:ns].to("cuda")
    bound = 20 * n * EPS
    # Q * L[:, cols] = A[:, cols]
    Lc = (torch.tril(F.factors)[:, cols].T.contiguous()).T
    QL = q.lmul_(F.Q, Lc)
    torch.cuda.synchronize()
    assert (torch.linalg.norm(QL - A[:, cols]) / torch.linalg.norm(A[:, cols])).item() <= bound
    # Q' Q e_j = e_j
    E = torch.zeros((ns, n), dtype=torch.float64, device="cuda").T
    E[cols, torch.arange(ns, device="cuda")] = 1.0
    E0 = E.clone()
    QE = q.lmul_(F.Q, E)
    torch.cuda.synchronize()
    assert abs(torch.linalg.norm(QE).item() - np.sqrt(ns)) <= bound * np.sqrt(ns)
    QtQE = q.lmul_(F.Q.T, QE)
    torch.cuda.synchronize()
    assert torch.linalg.norm(QtQE - E0).item() <= bound * np.sqrt(ns)


def test_config3_batched_262144_x_32x32_with_lmul(q, h):
    import torch

    batch, n, nrhs, ns = 262144, 32, 8, 256
    g = torch.Generator(device="cuda").manual_seed(1234321)
    A = torch.randn((batch, n, n), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((batch, nrhs, n), dtype=torch.float64, device="cuda", generator=g)
    F = A.clone()
    tau = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    q.ql_batched_(F, tau, handle=h)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(F).all()) and bool(torch.isfinite(tau).all())
    idx = torch.linspace(0, batch - 1, ns).long()
    Fo, tauo = cport.geql2_batched(A[idx.cuda()].cpu().numpy())
    nrm = np.linalg.norm(A[idx.cuda()].cpu().numpy().reshape(ns, -1), axis=1).max()
    assert np.abs(F[idx.cuda()].cpu().numpy() - Fo).max() <= 50 * n * EPS * nrm
    assert np.abs(tau[idx.cuda()].cpu().numpy() - tauo).max() <= 50 * n * EPS
    # L'L = A'A for every matrix (Q orthogonal): checked through the diagonal, ||a_j||^2 = ||L[:, j]||^2
    L = torch.triu(F)                # batched layout is F[b, col, row]: row >= col  <=>  upper triangle of the (col,row) view
    cn_A = (A * A).sum(dim=2)
    cn_L = (L * L).sum(dim=2)
    assert float(((cn_A - cn_L).abs() / cn_A).max()) <= 50 * n * EPS
    # lmul!(Q', B) preserves column norms, and lmul!(Q, .) undoes it -- for all 262144 matrices
    W = B.clone()
    q.lmul_batched_(F, tau, W, adjoint=True, handle=h)
    torch.cuda.synchronize()
    nb, nw = (B * B).sum(dim=2), (W * W).sum(dim=2)
    assert float(((nb - nw).abs() / nb).max()) <= 20 * n * EPS
    Wo = cport.lmul(np.asfortranarray(Fo[0].T), tauo[0], np.asfortranarray(B[idx[0]].cpu().numpy().T), adjoint=True)
    assert np.abs(W[idx[0]].cpu().numpy().T - Wo).max() <= 50 * n * EPS * np.abs(Wo).max() * n
    q.lmul_batched_(F, tau, W, adjoint=False, handle=h)
    torch.cuda.synchronize()
    assert float((W - B).abs().max()) <= 20 * n * EPS * float(B.abs().max()) * np.sqrt(n)


def test_config4_rq_tall_skinny_65536x1024(q, h):
    m, n = 65536, 1024
    rng = np.random.default_rng(1234321)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F, tau = q.rq_(A.copy(order="F"), handle=h)
    lapack.set_num_threads(16)
    Fo, tauo = lapack.gerqf(A)
    tol = 50 * m * EPS * np.linalg.norm(A)
    assert np.abs(tau - tauo).max() <= 50 * m * EPS
    assert np.abs(np.triu(F[m - n:, :]) - np.triu(Fo[m - n:, :])).max() <= tol
    assert np.abs(F - Fo).max() <= tol           # rows above the triangle hold the transformed rows of A


def test_config5_batched_1048576_x_16x16_f32_with_ldiv(q, h):
    import torch

    batch, n, nrhs, ns = 1048576, 16, 8, 256
    eps32 = float(np.finfo(np.float32).eps)
    g = torch.Generator(device="cuda").manual_seed(1234321)
    A = torch.randn((batch, n, n), dtype=torch.float32, device="cuda", generator=g)
    B = torch.randn((batch, nrhs, n), dtype=torch.float32, device="cuda", generator=g)
    F = A.clone()
    tau = torch.empty((batch, n), dtype=torch.float32, device="cuda")
    q.ql_batched_(F, tau, handle=h)
    X = B.clone()
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    q.ldiv_batched_(F, tau, X, info=info, handle=h)
    torch.cuda.synchronize()
    assert int(info.abs().max()) == 0
    idx = torch.linspace(0, batch - 1, ns).long().cuda()
    Fo, tauo = cport.geql2_batched(A[idx].cpu().numpy())
    assert np.abs(F[idx].cpu().numpy() - Fo).max() <= 50 * n * eps32 * np.linalg.norm(A[idx].cpu().numpy().reshape(ns, -1), axis=1).max()
    assert np.abs(tau[idx].cpu().numpy() - tauo).max() <= 50 * n * eps32
    # backward error of every solve, evaluated in Float64: ||A x - b|| <= c n eps ||A|| ||x||
    Ad = A.double().transpose(1, 2)                       # (batch, row, col)
    Xd = X.double().transpose(1, 2)                       # (batch, row, rhs)
    Bd = B.double().transpose(1, 2)
    R = torch.bmm(Ad, Xd) - Bd
    num = torch.linalg.norm(R.reshape(batch, -1), dim=1)
    den = torch.linalg.norm(Ad.reshape(batch, -1), dim=1) * torch.linalg.norm(Xd.reshape(batch, -1), dim=1)
    assert float((num / den).max()) <= 50 * n * eps32
