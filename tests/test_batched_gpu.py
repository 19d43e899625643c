"""GPU parity tests of the batched small-matrix kernels (K6) through the C-ABI vs the CPU oracle.

Tolerances are BASELINE.json's: L and tau elementwise within 50*n*eps*||A||; solves within the
reference's own 5000*eps-class tolerance (test/test_ql.jl:69) scaled by the solution size.
"""
import os

import numpy as np
import pytest

from oracle import cport, lapack
from __graft_entry__ import load_package

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ql_golden.npz"))


@pytest.fixture(scope="module")
def q():
    return load_package()


@pytest.fixture(scope="module")
def h(q):
    hd = q.Handle(0)
    yield hd
    hd.close()


def _eps(dt):
    return np.finfo(dt).eps


def _check_factors(F, tau, Fo, tauo, A, n, dt):
    nrm = np.linalg.norm(A.reshape(A.shape[0], -1), axis=1).max()
    t = 50 * n * _eps(dt) * max(nrm, 1.0)
    assert np.isfinite(F).all() and np.isfinite(tau).all()
    assert np.abs(F - Fo).max() <= t, f"factors differ by {np.abs(F - Fo).max():.3e} > {t:.3e}"
    assert np.abs(tau - tauo).max() <= 50 * n * _eps(dt), f"tau differs by {np.abs(tau - tauo).max():.3e}"


@pytest.mark.parametrize("n", [8, 16, 32])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("batch", [1, 3, 17, 1000, 4099])
def test_batched_geqlf_vs_oracle(q, h, n, dt, batch):
    rng = np.random.default_rng(n * 1000 + batch)
    A = rng.standard_normal((batch, n, n)).astype(dt)
    F, tau = q.ql_batched_(A.copy(), handle=h)
    Fo, tauo = cport.geql2_batched(A)
    _check_factors(F, tau, Fo, tauo, A, n, dt)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 9, 12, 15, 17, 20, 24, 31, 33, 40, 48, 56, 63, 64])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_batched_generic_sizes_vs_lapack(q, h, n, dt):
    """Orders other than 8/16/32: n < 32 runs on the next larger register tile (zero-padded in front, the padded steps are
    not taken), 32 < n <= 64 on the shared-memory kernels."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((70, n, n)).astype(dt)
    F, tau = q.ql_batched_(A.copy(), handle=h)
    Fo = A.copy()
    tauo = lapack.geqlf_loop(Fo)
    _check_factors(F, tau, Fo, tauo, A, n, dt)


@pytest.mark.parametrize("name,dt,n", [("b32_f64", np.float64, 32), ("b16_f32", np.float32, 16), ("b8_f64", np.float64, 8)])
def test_batched_vs_golden_lapack(q, h, name, dt, n):
    """Golden fixtures were produced by LAPACK ?geqlf (the routine behind ql!, src/ql.jl:72)."""
    A = GOLD[f"{name}_A"]
    F, tau = q.ql_batched_(A.copy(), handle=h)
    _check_factors(F, tau, GOLD[f"{name}_F"], GOLD[f"{name}_tau"], A, n, dt)
    B = GOLD[f"{name}_B"]
    tolb = 50 * n * _eps(dt) * np.abs(B).max() * n
    QtB = q.lmul_batched_(F, tau, B.copy(), adjoint=True, handle=h)
    assert np.abs(QtB - GOLD[f"{name}_QtB"]).max() <= tolb
    X = q.ldiv_batched_(F, tau, B.copy(), handle=h)
    Xo = GOLD[f"{name}_X"]
    # per-matrix tolerance scaled with the size of the solution (ill-conditioned draws happen)
    err = np.abs(X - Xo).reshape(X.shape[0], -1).max(axis=1)
    scale = np.abs(Xo).reshape(X.shape[0], -1).max(axis=1)
    cond = np.array([np.linalg.cond(A[i]) for i in range(A.shape[0])])
    assert (err <= 50 * n * _eps(dt) * cond * np.maximum(scale, 1.0)).all()


@pytest.mark.parametrize("n,dt", [(32, np.float64), (16, np.float32), (8, np.float64), (16, np.float64), (32, np.float32)])
@pytest.mark.parametrize("nrhs", [1, 2, 3, 8, 11])
def test_batched_lmul_and_solve_vs_oracle(q, h, n, dt, nrhs):
    rng = np.random.default_rng(7 * n + nrhs)
    batch = 37
    A = rng.standard_normal((batch, n, n)).astype(dt)
    Fo, tauo = cport.geql2_batched(A)
    B = rng.standard_normal((batch, nrhs, n)).astype(dt)
    tolb = 50 * n * _eps(dt) * np.abs(B).max() * np.sqrt(n)
    for adj in (False, True):
        got = q.lmul_batched_(Fo, tauo, B.copy(), adjoint=adj, handle=h)
        ref = np.stack([cport.lmul(Fo[i].T, tauo[i], B[i].T, adjoint=adj).T for i in range(batch)])
        assert np.abs(got - ref).max() <= tolb
    info = np.full(batch, -1, dtype=np.int32)
    X = q.ldiv_batched_(Fo, tauo, B.copy(), info=info, handle=h)
    assert (info == 0).all()
    Xo = np.stack([cport.ldiv(Fo[i].T, tauo[i], B[i].T).T for i in range(batch)])
    cond = np.array([np.linalg.cond(A[i].astype(np.float64)) for i in range(batch)])
    err = np.abs(X - Xo).reshape(batch, -1).max(axis=1)
    scale = np.maximum(np.abs(Xo).reshape(batch, -1).max(axis=1), 1.0)
    assert (err <= 50 * n * _eps(dt) * cond * scale).all()
    # residual: A x = b within the reference's solve tolerance class
    for i in range(batch):
        r = A[i].T.astype(np.float64) @ X[i].T.astype(np.float64) - B[i].T
        assert np.abs(r).max() <= 5000 * _eps(dt) * cond[i]


def test_batched_roundtrip_properties_full_size(q, h):
    """Size-independent properties at a large batch: ||Q'Q - I||, ||QL - A|| within 20*n*eps."""
    import torch

    n, batch = 32, 65536
    g = torch.Generator(device="cuda").manual_seed(1234321)
    A = torch.randn((batch, n, n), dtype=torch.float64, device="cuda", generator=g)
    F = A.clone()
    tau = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    q.ql_batched_(F, tau, handle=h)
    E = torch.eye(n, dtype=torch.float64, device="cuda").expand(batch, n, n).contiguous()
    q.lmul_batched_(F, tau, E, adjoint=False, handle=h)          # E_i = Q_i (stored column-major -> E[i] = Q_i^T)
    h.sync()
    Q = E.transpose(1, 2)
    L = torch.tril(F.transpose(1, 2))
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    orth = (Q.transpose(1, 2) @ Q - eye).flatten(1).norm(dim=1).max().item()
    Am = A.transpose(1, 2)
    rec = ((Q @ L - Am).flatten(1).norm(dim=1) / Am.flatten(1).norm(dim=1)).max().item()
    bound = 20 * n * np.finfo(np.float64).eps
    assert orth <= bound and rec <= bound, (orth, rec, bound)
    # run-to-run bitwise determinism (test/test_ql.jl:76 uses ==)
    F2 = A.clone()
    tau2 = torch.empty_like(tau)
    q.ql_batched_(F2, tau2, handle=h)
    h.sync()
    assert torch.equal(F, F2) and torch.equal(tau, tau2)


def test_batched_edge_conventions(q, h):
    """LAPACK ?larfg exits: zero tail -> tau = 0, beta = alpha; tiny/huge columns rescaled."""
    n = 32
    A = np.zeros((6, n, n))
    A[0] = np.eye(n)                                  # identity: all tails zero -> tau = 0, F = I
    A[1] = np.diag(np.arange(1, n + 1) * -1.0)        # negative diagonal stays negative (no sign flip)
    rng = np.random.default_rng(5)
    A[2] = rng.standard_normal((n, n)) * 1e-160       # squares underflow
    A[3] = rng.standard_normal((n, n)) * 1e160        # squares overflow
    A[4] = rng.standard_normal((n, n))
    A[4][:, 0] = 0.0                                  # a zero row of the (column-major) matrix
    A[5] = rng.standard_normal((n, n))
    A[5][3, :] = 0.0                                  # column 3 of the matrix entirely zero
    F, tau = q.ql_batched_(A.copy(), handle=h)
    Fo, tauo = cport.geql2_batched(A)
    assert np.array_equal(tau[0], np.zeros(n)) and np.array_equal(F[0], np.eye(n))
    assert np.array_equal(tau[1], np.zeros(n)) and np.array_equal(F[1], A[1])
    for i in (2, 3, 4, 5):
        sc = np.abs(A[i]).max()
        assert np.isfinite(F[i]).all()
        # column-major storage: F[i][col][row]; L lives at row >= col and scales with the data, the
        # reflector tails v (row < col) are scale free
        Lg, Lo = np.triu(F[i]), np.triu(Fo[i])
        assert np.abs(Lg - Lo).max() <= 50 * n * _eps(np.float64) * sc * n, i
        assert np.abs((F[i] - Lg) - (Fo[i] - Lo)).max() <= 50 * n * _eps(np.float64), i
        assert np.abs(tau[i] - tauo[i]).max() <= 50 * n * _eps(np.float64), i


def test_batched_strided_layout_and_bitwise_lda_independence(q, h):
    """test/test_ql.jl:76: the result must not depend on the leading dimension (bitwise)."""
    n, batch, lda = 16, 50, 18
    rng = np.random.default_rng(11)
    A = rng.standard_normal((batch, n, n))
    F, tau = q.ql_batched_(A.copy(), handle=h)
    P = np.zeros((batch, n + 1, lda))
    P[:, :n, :n] = A
    taup = np.zeros((batch, n + 3))
    rc = h.call("qlb200_dgeqlf_batched", n, P.ctypes.data, lda, (n + 1) * lda, taup.ctypes.data, n + 3, batch)
    assert rc == 0
    assert np.array_equal(P[:, :n, :n], F) and np.array_equal(taup[:, :n], tau)
    assert (P[:, n, :] == 0).all() and (P[:, :, n:] == 0).all()      # padding untouched


def test_batched_argument_errors(q, h):
    A = np.zeros((2, 65, 65))
    with pytest.raises(q.ArgumentError):
        q.ql_batched_(A, handle=h)                         # n > 64 unsupported
    with pytest.raises(q.ArgumentError):
        q.lmul_batched_(np.zeros((2, 65, 65)), np.zeros((2, 65)), np.zeros((2, 1, 65)), handle=h)   # apply: n <= 64 too
    A = np.zeros((0, 32, 32))
    F, tau = q.ql_batched_(A, handle=h)                    # empty batch is a no-op
    assert tau.shape == (0, 32)
    with pytest.raises(q.DimensionMismatch):
        q.lmul_batched_(np.zeros((2, 32, 32)), np.zeros((2, 32)), np.zeros((2, 4, 16)), handle=h)


@pytest.mark.parametrize("n", [1, 2, 5, 7, 9, 12, 20, 24, 31, 33, 48, 64])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("nrhs", [1, 3, 8, 11])
def test_batched_generic_sizes_lmul_and_solve_vs_oracle(q, h, n, dt, nrhs):
    """Sizes without a register apply kernel go through the generic warp-per-matrix kernel (any n <= 64)."""
    rng = np.random.default_rng(100 * n + nrhs)
    batch = 23
    A = rng.standard_normal((batch, n, n)).astype(dt)
    Fo, tauo = cport.geql2_batched(A)
    B = rng.standard_normal((batch, nrhs, n)).astype(dt)
    tolb = 50 * n * _eps(dt) * np.abs(B).max() * np.sqrt(n)
    for adj in (False, True):
        got = q.lmul_batched_(Fo, tauo, B.copy(), adjoint=adj, handle=h)
        ref = np.stack([cport.lmul(Fo[i].T, tauo[i], B[i].T, adjoint=adj).T for i in range(batch)])
        assert np.abs(got - ref).max() <= tolb
    info = np.full(batch, -1, dtype=np.int32)
    X = q.ldiv_batched_(Fo, tauo, B.copy(), info=info, handle=h)
    assert (info == 0).all()
    Xo = np.stack([cport.ldiv(Fo[i].T, tauo[i], B[i].T).T for i in range(batch)])
    cond = np.array([np.linalg.cond(A[i].astype(np.float64)) for i in range(batch)])
    err = np.abs(X - Xo).reshape(batch, -1).max(axis=1)
    scale = np.maximum(np.abs(Xo).reshape(batch, -1).max(axis=1), 1.0)
    assert (err <= 50 * n * _eps(dt) * cond * scale).all()


def test_batched_generic_singular_info_and_strides(q, h):
    n = 12
    A = np.random.default_rng(3).standard_normal((4, n, n))
    Fo, tauo = cport.geql2_batched(A)
    Fo[2, 4, 4] = 0.0                                      # L[5,5] = 0 in matrix 2
    Fo[2, 9, 9] = 0.0                                      # ... and L[10,10]: the FIRST zero is reported
    info = np.zeros(4, dtype=np.int32)
    q.ldiv_batched_(Fo, tauo, np.ones((4, 1, n)), info=info, handle=h)
    assert list(info) == [0, 0, 5, 0]


def test_batched_singular_info(q, h):
    n = 8
    A = np.random.default_rng(3).standard_normal((4, n, n))
    Fo, tauo = cport.geql2_batched(A)
    Fo[2, 4, 4] = 0.0                                      # L[5,5] = 0 in matrix 2
    info = np.zeros(4, dtype=np.int32)
    q.ldiv_batched_(Fo, tauo, np.ones((4, 1, n)), info=info, handle=h)
    assert list(info) == [0, 0, 5, 0]


@pytest.mark.parametrize("n,dt", [(32, np.float64), (16, np.float32), (8, np.float64)])
def test_batched_unaligned_layouts_take_the_plain_copy_path(q, h, n, dt):
    """Odd leading dimensions / strides (not multiples of 16 bytes) cannot use TMA: the kernels fall back to plain loads
    and stores and must give bitwise the same factors and products as the aligned layout."""
    batch, lda, ldb, nrhs = 21, n + 1, n + 3, 5
    rng = np.random.default_rng(n)
    A = rng.standard_normal((batch, n, n)).astype(dt)
    B = rng.standard_normal((batch, nrhs, n)).astype(dt)
    F, tau = q.ql_batched_(A.copy(), handle=h)
    W = q.lmul_batched_(F, tau, B.copy(), adjoint=True, handle=h)
    X = q.ldiv_batched_(F, tau, B.copy(), handle=h)
    sA, sT, sB = n * lda + 1, n + 1, nrhs * ldb + 1
    P = np.zeros(batch * sA + 8, dtype=dt)
    T = np.zeros(batch * sT + 8, dtype=dt)
    PB = np.zeros(batch * sB + 8, dtype=dt)
    for b in range(batch):
        P[b * sA:b * sA + n * lda].reshape(n, lda)[:, :n] = A[b]
        PB[b * sB:b * sB + nrhs * ldb].reshape(nrhs, ldb)[:, :n] = B[b]
    pfx = "d" if dt == np.float64 else "s"
    assert h.call(f"qlb200_{pfx}geqlf_batched", n, P.ctypes.data, lda, sA, T.ctypes.data, sT, batch) == 0
    for b in range(batch):
        assert np.array_equal(P[b * sA:b * sA + n * lda].reshape(n, lda)[:, :n], F[b]), b
        assert np.array_equal(T[b * sT:b * sT + n], tau[b]), b
    PB2 = PB.copy()
    assert h.call(f"qlb200_{pfx}ormql_batched", b"T", n, nrhs, P.ctypes.data, lda, sA, T.ctypes.data, sT, PB.ctypes.data, ldb, sB, batch) == 0
    assert h.call(f"qlb200_{pfx}qlsolve_batched", n, nrhs, P.ctypes.data, lda, sA, T.ctypes.data, sT, PB2.ctypes.data, ldb, sB, batch, 0) == 0
    for b in range(batch):
        assert np.array_equal(PB[b * sB:b * sB + nrhs * ldb].reshape(nrhs, ldb)[:, :n], W[b]), b
        assert np.array_equal(PB2[b * sB:b * sB + nrhs * ldb].reshape(nrhs, ldb)[:, :n], X[b]), b


@pytest.mark.parametrize("n,dt", [(32, np.float64), (16, np.float32), (24, np.float64)])
def test_batched_host_chunk_pipeline_same_bits(q, h, n, dt):
    """Host-pointer calls run as an upload | compute | download pipeline over chunks of matrices ("batch_chunk_mb");
    the chunking must not change a bit (chunk boundaries that split the batch unevenly, ldiv info included)."""
    rng = np.random.default_rng(n)
    batch = 3001
    A = rng.standard_normal((batch, n, n)).astype(dt)
    B = rng.standard_normal((batch, 5, n)).astype(dt)
    A[1234, 3, 3] = 0.0
    out = {}
    try:
        for mb in (0, 1):
            h.set_option("batch_chunk_mb", mb)
            F, tau = q.ql_batched_(A.copy(), handle=h)
            W = q.lmul_batched_(F, tau, B.copy(), adjoint=True, handle=h)
            Fz = F.copy()
            Fz[77, 2, 2] = 0.0                                 # zero on the diagonal of L in matrix 77
            info = np.full(batch, -1, dtype=np.int32)
            X = q.ldiv_batched_(Fz, tau, B.copy(), info=info, handle=h)
            out[mb] = (F, tau, W, X, info)
    finally:
        h.set_option("batch_chunk_mb", 32)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b, equal_nan=True)
    assert out[1][4][77] == 3 and (np.delete(out[1][4], 77) == 0).all()


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("bvar", [2, 3])
def test_batched_n32_single_and_split_kernels_vs_oracle(q, h, dt, bvar):
    """n = 32 has two code paths (one kernel for small batches, the two-phase split for large ones); both are forced here
    at a small batch, including matrices that the register kernels flag and leave to the generic kernel."""
    rng = np.random.default_rng(bvar)
    batch = 1537
    A = rng.standard_normal((batch, 32, 32)).astype(dt)
    tiny = 1e-200 if dt == np.float64 else 1e-30
    A[5] *= tiny                                   # flagged in phase A (squares underflow)
    A[9, :16, :16] *= tiny                         # leading block: flagged in phase B only
    A[11, :, :8] = 0.0                             # zero rows: tau = 0 steps
    try:
        h.set_option("bvar", bvar)
        F, tau = q.ql_batched_(A.copy(), handle=h)
    finally:
        h.set_option("bvar", 0)
    Fo = A.copy()
    tauo = lapack.geqlf_loop(Fo)
    eps = _eps(dt)
    for i in (5, 9, 11, 0, 1, batch - 1):
        # layout F[b, col, row]: L (row >= col) scales with ||A||, the reflector tails (row < col) are O(1) whatever the scale
        amax = float(np.abs(A[i]).max())                        # scaled norm: the squares of matrix 5 underflow
        nrm = amax * np.linalg.norm(A[i].astype(np.float64) / amax)
        d = np.abs(F[i].astype(np.float64) - Fo[i].astype(np.float64))
        assert np.triu(d).max() <= 50 * 32 * eps * nrm, i
        assert np.tril(d, -1).max() <= 50 * 32 * eps, i
    _check_factors(F, tau, Fo, tauo, A, 32, dt)


@pytest.mark.parametrize("n,dt", [(32, np.float64), (16, np.float32), (24, np.float64), (48, np.float64)])
@pytest.mark.parametrize("op", ["T", "N", "S"])
@pytest.mark.parametrize("where", ["host", "device"])
def test_fused_factor_apply_same_bits_as_two_calls(q, h, n, dt, op, where):
    """qlb200_?geqlf_apply_batched (BASELINE configs[2]/[4] as one call) = ql_batched_ followed by lmul_batched_ / ldiv_batched_:
    same kernels, so the same bits -- across several L2 chunks (fuse_chunk) and host pipeline chunks."""
    import torch

    rng = np.random.default_rng(n + ord(op))
    batch, nrhs = 1500, 8
    A = rng.standard_normal((batch, n, n)).astype(dt)
    B = rng.standard_normal((batch, nrhs, n)).astype(dt)
    F2, t2 = q.ql_batched_(A.copy(), handle=h)
    if op == "S":
        X2 = q.ldiv_batched_(F2, t2, B.copy(), handle=h)
    else:
        X2 = q.lmul_batched_(F2, t2, B.copy(), adjoint=(op == "T"), handle=h)
    h.set_option("fuse_chunk", 400)
    h.set_option("batch_chunk_mb", 1)
    try:
        info = np.full(batch, -1, dtype=np.int32)
        if where == "host":
            F1, t1, X1 = q.ql_apply_batched_(A.copy(), B.copy(), op=op, info=info, handle=h)
        else:
            Ad, Bd = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
            idev = torch.from_numpy(info).cuda()
            F1, t1, X1 = q.ql_apply_batched_(Ad, Bd, op=op, info=idev, handle=h)
            torch.cuda.synchronize()
            F1, t1, X1, info = F1.cpu().numpy(), t1.cpu().numpy(), X1.cpu().numpy(), idev.cpu().numpy()
    finally:
        h.set_option("fuse_chunk", 0)
        h.set_option("batch_chunk_mb", 32)
    assert np.array_equal(F1, F2) and np.array_equal(t1, t2) and np.array_equal(X1, X2)
    if op == "S":
        assert (info == 0).all()


@pytest.mark.parametrize("n", [33, 40, 47, 48, 56, 64])
@pytest.mark.parametrize("bvar", [6, 7])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_batched_wy_pair_and_single_warp_kernels_same_bits(q, h, n, bvar, dt):
    """32 < n <= 64: the warp-pair kernel (panel warp + update warp, look-ahead inside the matrix; default for n > 56) and the
    one-warp-per-matrix kernel execute the same arithmetic per matrix: identical bits, and both match LAPACK."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((150, n, n)).astype(dt)
    F0, t0 = q.ql_batched_(A.copy(), handle=h)
    h.set_option("bvar", bvar)
    try:
        F1, t1 = q.ql_batched_(A.copy(), handle=h)
    finally:
        h.set_option("bvar", 0)
    assert np.array_equal(F0, F1) and np.array_equal(t0, t1)
    Fo = A.copy()
    tauo = lapack.geqlf_loop(Fo)
    _check_factors(F1, t1, Fo, tauo, A, n, dt)
