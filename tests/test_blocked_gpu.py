"""GPU parity tests of the single-matrix path (blocked ql!, lmul!/rmul!, ldiv!, rq!) through the
C-ABI vs the oracle: LAPACK ?geqlf/?gerqf (the routines behind the reference's ql!/rq!,
src/ql.jl:72, src/rq.jl:401) and the C restatement of the reference's scalar loops.

Tolerances (BASELINE.json north_star): L and tau elementwise within 50*n*eps*||A||;
||QL - A||/||A|| and ||Q'Q - I|| within 20*n*eps; solves like the reference's own test
(test/test_ql.jl:69: 5000*eps class, scaled by cond).
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


def _tol(shape, A):
    return 50 * max(shape) * EPS * np.linalg.norm(A)


@pytest.mark.parametrize("transa,transb", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")])
@pytest.mark.parametrize("mnk", [(128, 128, 128), (300, 200, 100), (257, 129, 77), (1000, 16, 500), (64, 9, 1000), (5, 3, 2)])
@pytest.mark.parametrize("cfg", [-1, 0, 1, 3])
def test_dmma_gemm_vs_torch(q, h, transa, transb, mnk, cfg):
    import torch

    m, n, k = mnk
    g = torch.Generator(device="cuda").manual_seed(m * 7 + n)
    cm = lambda r, c: torch.randn((c, r), dtype=torch.float64, device="cuda", generator=g).T   # column-major r x c
    A = cm(k, m) if transa == "T" else cm(m, k)
    B = cm(n, k) if transb == "T" else cm(k, n)
    C = cm(m, n)
    ref = 1.5 * (A.T if transa == "T" else A) @ (B.T if transb == "T" else B) - 0.5 * C
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    h.set_option("gemm_cfg", cfg)
    try:
        h.call("qlb200_dgemm_dev", transa.encode(), transb.encode(), m, n, k, 1.5, A.data_ptr(), A.stride(1),
               B.data_ptr(), B.stride(1), -0.5, C.data_ptr(), C.stride(1))
    finally:
        h.set_option("gemm_cfg", -1)
    torch.cuda.synchronize()
    assert (C - ref).abs().max().item() <= 20 * k * EPS * 10


SHAPES = [(10, 10), (10, 12), (12, 10), (3, 5), (1, 1), (5, 1), (1, 5), (16, 16), (17, 17), (100, 100), (100, 103),
          (129, 129), (200, 150), (150, 420), (300, 300), (513, 257), (1024, 1024), (1500, 700), (2100, 2100)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("lookahead", [0, 1])
def test_geqlf_matches_lapack(q, h, shape, lookahead):
    """test/test_ql.jl:5-33 at the LAPACK boundary: same packed factors and tau."""
    A = np.asfortranarray(np.random.default_rng(shape[0] * 31 + shape[1]).standard_normal(shape) / 2)
    h.set_option("lookahead", lookahead)
    try:
        F = q.ql(A, handle=h)
    finally:
        h.set_option("lookahead", 0)
    Fo, tauo = lapack.geqlf(A)
    t = _tol(shape, A)
    assert np.isfinite(F.factors).all()
    assert np.abs(F.factors - Fo).max() <= t, np.abs(F.factors - Fo).max()
    assert np.abs(F.tau - tauo).max() <= t


@pytest.mark.parametrize("nb", [16, 32, 64, 96, 128, 176, 256])
@pytest.mark.parametrize("lookahead,rpt", [(0, 4), (1, 4), (1, 1), (0, 2)])
def test_geqlf_panel_widths(q, h, nb, lookahead, rpt):
    A = np.asfortranarray(np.random.default_rng(nb).standard_normal((700, 450)))
    h.set_option("nb", nb)
    h.set_option("lookahead", lookahead)
    h.set_option("strip_rpt", rpt)
    try:
        F = q.ql(A, handle=h)
    finally:
        h.set_option("nb", 0)
        h.set_option("lookahead", 0)
        h.set_option("strip_rpt", 4)
    Fo, tauo = lapack.geqlf(A)
    assert np.abs(F.factors - Fo).max() <= _tol(A.shape, A) and np.abs(F.tau - tauo).max() <= _tol(A.shape, A)


def test_geqlf_golden_fixtures(q, h):
    for name in ["sq10", "wide10x12", "tall12x10", "wide3x5", "sq100", "wide100x103", "sq200", "tall300x130"]:
        A = GOLD[f"{name}_A"]
        F = q.ql(A, handle=h)
        t = _tol(A.shape, A)
        assert np.abs(F.factors - GOLD[f"{name}_F"]).max() <= t, name
        assert np.abs(F.tau - GOLD[f"{name}_tau"]).max() <= t, name


def test_issue_7304_sign(q, h):
    """test/test_ql.jl:120-124: A = [-s -s; -s s] => Q == -A."""
    A = GOLD["i7304_A"]
    F = q.ql(A, handle=h)
    Q = F.Q.matrix()
    assert np.linalg.norm(A + Q) < 4 * EPS


def test_lda_padding_and_bitwise_lda_independence(q, h):
    """test/test_ql.jl:76: results must not depend on the leading dimension (==), padding untouched."""
    rng = np.random.default_rng(9)
    A = np.asfortranarray(rng.standard_normal((200, 90)))
    F1 = q.ql(A, handle=h)
    big = np.full((211, 95), 7.0, order="F")
    big[:200, :90] = A
    view = big[:200, :90]
    F2 = q.ql_(view, handle=h)
    assert np.array_equal(F1.factors, F2.factors) and np.array_equal(F1.tau, F2.tau)
    assert (big[200:, :] == 7.0).all() and (big[:, 90:] == 7.0).all()
    F3 = q.ql(A, handle=h)
    assert np.array_equal(F1.factors, F3.factors)          # run-to-run determinism


def test_zero_columns_follow_lapack(q, h):
    """LAPACK ?larfg: an exactly-zero tail gives tau = 0 and keeps the sign."""
    A = np.asfortranarray(np.eye(40))
    F = q.ql(A, handle=h)
    assert np.array_equal(F.tau, np.zeros(40)) and np.array_equal(F.factors, np.eye(40))
    A = np.asfortranarray(np.random.default_rng(2).standard_normal((60, 60)))
    A[:, 17] = 0.0
    F = q.ql(A, handle=h)
    Fo, tauo = lapack.geqlf(A)
    assert np.abs(F.factors - Fo).max() <= _tol(A.shape, A) and np.abs(F.tau - tauo).max() <= _tol(A.shape, A)


@pytest.mark.parametrize("shape", [(100, 100), (100, 103), (300, 300), (500, 200)])
@pytest.mark.parametrize("nrhs", [1, 2, 33])
def test_lmul_rmul_vs_reference_loops(q, h, shape, nrhs):
    """test/test_ql.jl:157-181: lmul!/rmul! with Q and Q' vs the reference's scalar loops."""
    rng = np.random.default_rng(5)
    A = np.asfortranarray(rng.standard_normal(shape))
    Fo, tauo = lapack.geqlf(A)
    F = q.QL(Fo, tauo, handle=h)
    m = shape[0]
    B = np.asfortranarray(rng.standard_normal((m, nrhs)))
    C = np.asfortranarray(rng.standard_normal((nrhs, m)))
    t = 50 * m * EPS * 10
    for adj in (False, True):
        Qx = F.Q.T if adj else F.Q
        assert np.abs(q.lmul_(Qx, B.copy(order="F")) - cport.lmul(Fo, tauo, B, adjoint=adj)).max() <= t
        assert np.abs(q.rmul_(C.copy(order="F"), Qx) - cport.rmul(C, Fo, tauo, adjoint=adj)).max() <= t
    with pytest.raises(q.DimensionMismatch):
        q.lmul_(F.Q, np.zeros((m + 1, 2), order="F"))          # src/ql.jl:326-328
    with pytest.raises(q.DimensionMismatch):
        q.rmul_(np.zeros((2, m + 1), order="F"), F.Q)          # src/ql.jl:385-387


@pytest.mark.parametrize("n", [10, 100, 257, 1024])
def test_orthogonality_reconstruction_and_solve(q, h, n):
    """test/test_ql.jl:63-70: Q'Q = I, Q L = A, a*(F\\b) = b."""
    rng = np.random.default_rng(n)
    A = np.asfortranarray(rng.standard_normal((n, n)) / 2)
    F = q.ql(A, handle=h)
    Q = F.Q.matrix()
    L = F.L
    bound = 20 * n * EPS
    assert np.linalg.norm(Q.T @ Q - np.eye(n)) <= bound * np.sqrt(n)
    assert np.linalg.norm(Q @ L - A) / np.linalg.norm(A) <= bound
    b = np.asfortranarray(rng.standard_normal((n, 3)))
    x = F.solve(b)
    Fo, tauo = lapack.geqlf(A)
    xo = cport.ldiv(Fo, tauo, b)
    cond = np.linalg.cond(A)
    assert np.abs(x - xo).max() <= 50 * n * EPS * cond * max(1.0, np.abs(xo).max())
    assert np.abs(A @ x - b).max() <= 5000 * EPS * cond
    xv = F.solve(b[:, 0].copy())
    assert xv.shape == (n,) and np.abs(xv - x[:, 0]).max() <= 50 * n * EPS * cond * max(1.0, np.abs(xo).max())


def test_solve_singular_and_dimension_errors(q, h):
    A = np.asfortranarray(np.random.default_rng(1).standard_normal((20, 20)))
    Fo, tauo = lapack.geqlf(A)
    Fo[7, 7] = 0.0
    with pytest.raises(q.SingularException) as e:
        q.ldiv_(q.QL(Fo, tauo, handle=h), np.ones((20, 1), order="F"))
    assert e.value.info == 8
    with pytest.raises(q.DimensionMismatch):
        q.ldiv_(q.QL(Fo, tauo, handle=h), np.ones((19, 1), order="F"))


@pytest.mark.parametrize("shape", [(10, 10), (8, 13), (40, 12), (300, 300), (130, 700), (900, 64)])
def test_rq_matches_gerqf(q, h, shape):
    """test/test_rq.jl:29-67: rq! vs LAPACK gerqf directly."""
    A = np.asfortranarray(np.random.default_rng(sum(shape)).standard_normal(shape) / 2)
    F, tau = q.rq_(A.copy(order="F"), handle=h)
    Fo, tauo = lapack.gerqf(A)
    t = _tol(shape, A)
    assert np.abs(F - Fo).max() <= t and np.abs(tau - tauo).max() <= t
    for name in ("rq_sq10", "rq_wide8x13", "rq_tall40x12"):
        G, gt = q.rq_(np.asfortranarray(GOLD[f"{name}_A"]).copy(order="F"), handle=h)
        assert np.abs(G - GOLD[f"{name}_F"]).max() <= 1e-12 and np.abs(gt - GOLD[f"{name}_tau"]).max() <= 1e-12


def test_device_pointer_path_large_properties(q, h):
    """Full-size style check through the _dev entry points on torch tensors: n = 4096,
    ||Q'Q - I|| and ||QL - A||/||A|| within 20*n*eps, run-to-run bitwise determinism."""
    import torch

    n = 4096
    g = torch.Generator(device="cuda").manual_seed(1234321)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T      # column-major
    Fm = A.T.contiguous().T.clone()
    Fm = (A.T.clone()).T
    F = q.ql_(Fm, handle=h)
    torch.cuda.synchronize()
    Fm2 = (A.T.clone()).T
    F2 = q.ql_(Fm2, handle=h)
    torch.cuda.synchronize()
    assert torch.equal(F.factors, F2.factors) and torch.equal(F.tau, F2.tau)
    Q = F.Q.matrix()
    torch.cuda.synchronize()
    L = F.L
    bound = 20 * n * EPS
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert torch.linalg.norm(Q.T @ Q - eye).item() <= bound * np.sqrt(n)
    assert (torch.linalg.norm(Q @ L - A) / torch.linalg.norm(A)).item() <= bound
    # flip identity (test/test_ql.jl:7-13): R of QR of the doubly flipped matrix is L flipped
    R = torch.linalg.qr(torch.flip(A, dims=(0, 1)), mode="r").R
    Lf = torch.flip(L, dims=(0, 1))
    sgn = torch.sign(torch.diagonal(R)) * torch.sign(torch.diagonal(Lf))
    assert (torch.abs(R * sgn[:, None] - Lf).max() / torch.abs(L).max()).item() <= 50 * n * EPS * 10


def test_t_cache_hits_and_misses(q, h):
    """ql! keeps the panels' T factors; lmul!/ldiv! reuse them only for the factors they belong to (device: same
    pointers; host: same fingerprint).  A stale cache must never be used: interleave two factorizations."""
    import torch

    rng = np.random.default_rng(77)
    n = 300
    A1 = np.asfortranarray(rng.standard_normal((n, n)))
    A2 = np.asfortranarray(rng.standard_normal((n, n)))
    B = np.asfortranarray(rng.standard_normal((n, 3)))
    F1 = q.ql(A1, handle=h)
    F2 = q.ql(A2, handle=h)                      # the cache now belongs to A2
    tol = 50 * n * EPS * np.abs(B).max() * n
    for F in (F1, F2, F1):                       # miss (rebuild), hit, miss
        got = q.lmul_(F.Q.T, B.copy(order="F"))
        ref = cport.lmul(F.factors, F.tau, B, adjoint=True)
        assert np.abs(got - ref).max() <= tol
    x = F2.solve(B)                              # hit
    assert np.abs(x - cport.ldiv(F2.factors, F2.tau, B)).max() <= 50 * n * EPS * np.linalg.cond(A2) * max(1.0, np.abs(x).max())
    # device pointers: cached and rebuilt T give the same answer
    g = torch.Generator(device="cuda").manual_seed(3)
    Ad = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T
    Bd = torch.randn((4, n), dtype=torch.float64, device="cuda", generator=g).T
    Fd = q.ql_((Ad.T.clone()).T, handle=h)
    r_hit = q.lmul_(Fd.Q, (Bd.T.clone()).T)
    torch.cuda.synchronize()
    h.set_option("tcache", 0)
    try:
        r_miss = q.lmul_(Fd.Q, (Bd.T.clone()).T)
        torch.cuda.synchronize()
    finally:
        h.set_option("tcache", 1)
    assert torch.equal(r_hit, r_miss)


@pytest.mark.parametrize("shape", [(20000, 40), (17000, 700)])
def test_geqlf_taller_than_the_cluster_capacity(q, h, shape):
    """m > 16384: strips that do not fit one cluster take the global-memory panel path (mixed with the cluster kernel
    once the strips get short enough); same LAPACK parity bar."""
    m, n = shape
    rng = np.random.default_rng(m)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F = q.ql(A, handle=h)
    lapack.set_num_threads(16)
    Fo, tauo = lapack.geqlf(A)
    tol = 50 * m * EPS * np.linalg.norm(A)
    assert np.abs(F.tau - tauo).max() <= 50 * m * EPS
    assert np.abs(F.factors - Fo).max() <= tol
    B = np.asfortranarray(rng.standard_normal((m, 2)))
    got = q.lmul_(F.Q.T, B.copy(order="F"))
    ref = lapack.ormql("L", "T", np.asfortranarray(Fo), tauo, B.copy(order="F"))
    assert np.abs(got - ref).max() <= 50 * m * EPS * np.abs(B).max() * np.sqrt(m)


@pytest.mark.parametrize("shape", [(2500, 2500), (2100, 2600), (3000, 2100), (4096, 4096)])
def test_host_pointer_ql_streams_in_and_out(q, h, shape):
    """Host-pointer ql! of a large matrix: the matrix streams in by column chunks while the right-hand panels are being
    factored (chunks join the trailing update on a fixed schedule) and finished panels stream back.  Same parity bar as
    the resident path, and the same answer as with streaming switched off."""
    m, n = shape
    rng = np.random.default_rng(m + n)
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F = q.ql(A, handle=h)
    lapack.set_num_threads(16)
    Fo, tauo = lapack.geqlf(A)
    tol = 50 * max(m, n) * EPS * np.linalg.norm(A)
    assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * EPS
    assert np.abs(F.factors - Fo).max() <= tol
    h.set_option("stream_in", 0)
    try:
        F0 = q.ql(A, handle=h)
    finally:
        h.set_option("stream_in", 1)
    assert np.abs(F.factors - F0.factors).max() <= tol and np.abs(F.tau - F0.tau).max() <= 50 * max(m, n) * EPS
    # the T cache filled by the streamed factorization serves the solve that follows (fingerprint hit)
    if m == n:
        b = np.asfortranarray(rng.standard_normal((n, 2)))
        x = F.solve(b)
        assert np.abs(A @ x - b).max() <= 5000 * EPS * np.linalg.cond(A)


@pytest.mark.parametrize("shape,split_min", [((2500, 2500), 1024), ((3000, 2100), 1024), ((5000, 4600), 1024), ((8192, 8192), 8192)])
def test_host_pointer_ql_split_schedule(q, h, shape, split_min):
    """Host-pointer ql! of a large tall/square matrix, split schedule (capi.cu): the right quarter of the columns is uploaded
    and factored while the rest is on its way, Q_R' is applied to the left part, then the leading block is factored; finished
    panels stream home.  Same arithmetic as one ql! (column blocks, right to left): the parity bar against LAPACK dgeqlf,
    agreement with the chunk-joining schedule (host_split = 0), nothing outside the matrix touched (lda > m)."""
    m, n = shape
    rng = np.random.default_rng(m + n)
    lda = m + 3
    buf = np.zeros((lda, n), order="F")
    buf[:m, :] = rng.standard_normal((m, n))
    A = buf[:m, :]
    h.set_option("host_split_min", split_min)
    h.set_option("host_split", 1)
    try:
        F = q.ql_(np.asfortranarray(A.copy()), handle=h)
        Fv = buf.copy(order="F")
        Av = Fv[:m, :]
        q.ql_(Av, handle=h)                                   # a view with lda > m
    finally:
        h.set_option("host_split_min", 8192)
    lapack.set_num_threads(16)
    Fo, tauo = lapack.geqlf(np.asfortranarray(A.copy()))
    tol = 50 * max(m, n) * EPS * np.linalg.norm(A)
    assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * EPS
    assert np.abs(F.factors - Fo).max() <= tol
    assert np.array_equal(Av, F.factors) and np.all(Fv[m:, :] == 0.0)
    h.set_option("host_split", 0)
    try:
        F0 = q.ql(np.asfortranarray(A.copy()), handle=h)
    finally:
        h.set_option("host_split", 2)
    assert np.abs(F.factors - F0.factors).max() <= tol and np.abs(F.tau - F0.tau).max() <= 50 * max(m, n) * EPS


@pytest.mark.parametrize("shape,cw,ramp,jc,pn", [((2500, 2500), 512, 4, 2, 8192), ((3000, 2100), 1024, 4, 2, 8192), ((5000, 4600), 512, 2, 3, 8192),
                                                 ((4096, 4096), 1024, 64, 1, 8192), ((8192, 8192), 1024, 4, 2, 8192),
                                                 ((2500, 2500), 512, 4, 2, 1), ((5000, 4600), 512, 2, 3, 2048), ((4096, 4096), 1024, 4, 2, 1024)])
def test_host_pointer_ql_chunks_join_the_lookahead(q, h, shape, cw, ramp, jc, pn):
    """Default host-pointer schedule for large tall/square matrices (host_split = 2): column chunks are uploaded right to left
    and JOIN the look-ahead factorization (blocked.cu) -- a late chunk first catches up with the panels factored so far (clean V
    from the packed factors, cached T) and then takes part in the trailing updates; finished panels stream home.  Per column the
    same sequence of block reflectors as LAPACK dgeqlf, so: the parity bar against dgeqlf, agreement with the schedule without
    look-ahead, lda > m respected, and lmul! right after (validated T cache of the streamed factorization).  pn =
    "la_pnarrow_min": with the next-panel update on the panel partition, that partition also has to wait for the joins."""
    m, n = shape
    rng = np.random.default_rng(m + 3 * n)
    lda = m + 5
    buf = np.zeros((lda, n), order="F")
    buf[:m, :] = rng.standard_normal((m, n))
    A = buf[:m, :]
    for name, v in (("host_split_min", 1024), ("la_min", 1024), ("chunk_cols", cw), ("la_join_ramp", ramp), ("la_join_chunks", jc),
                    ("host_split", 2), ("la_pnarrow_min", pn)):
        h.set_option(name, v)
    try:
        F = q.ql_(np.asfortranarray(A.copy()), handle=h)
        assert h.upload_rate() > 0.1
        B = np.asfortranarray(rng.standard_normal((m, 3)))
        QtB = q.lmul_(F.Q.T, B.copy(order="F"))
        Fv = buf.copy(order="F")
        Av = Fv[:m, :]
        q.ql_(Av, handle=h)                                   # a view with lda > m
        h.set_option("host_split", 0)
        F0 = q.ql(np.asfortranarray(A.copy()), handle=h)
    finally:
        for name, v in (("host_split_min", 8192), ("la_min", 4096), ("chunk_cols", 1024), ("la_join_ramp", 4), ("la_join_chunks", 2),
                        ("host_split", 2), ("la_pnarrow_min", 8192)):
            h.set_option(name, v)
    lapack.set_num_threads(16)
    Fo, tauo = lapack.geqlf(np.asfortranarray(A.copy()))
    tol = 50 * max(m, n) * EPS * np.linalg.norm(A)
    assert np.abs(F.tau - tauo).max() <= 50 * max(m, n) * EPS
    assert np.abs(F.factors - Fo).max() <= tol
    assert np.array_equal(Av, F.factors) and np.all(Fv[m:, :] == 0.0)
    assert np.abs(F.factors - F0.factors).max() <= tol and np.abs(F.tau - F0.tau).max() <= 50 * max(m, n) * EPS
    ref = lapack.ormql("L", "T", Fo, tauo, B.copy(order="F"))
    assert np.abs(QtB - ref).max() <= 50 * max(m, n) * EPS * np.linalg.norm(B)


def test_host_pointer_ql_pinned_matrix_replays_as_a_graph(q, h):
    """The same large host-pointer ql! again on the same PINNED matrix: first call plain launches, second call captured
    (uploads, both SM partitions' kernels, downloads) and launched as one CUDA graph, then replays.  Every call must leave the
    same bits (the joining schedule is fixed, not timing-driven) and follow new CONTENT in the same buffer; a pageable matrix
    never takes the graph and gives the same bits too."""
    import torch
    n = 2304
    rng = np.random.default_rng(5)
    hp = torch.empty((n, n), dtype=torch.float64, pin_memory=True)          # column-major n x n through the transpose view
    Ah = hp.numpy().T
    A1, A2 = rng.standard_normal((n, n)), rng.standard_normal((n, n))
    for name, v in (("host_split_min", 1024), ("la_min", 1024), ("chunk_cols", 512)):
        h.set_option(name, v)
    try:
        outs = []
        for A in (A1, A1, A1, A2, A1):
            Ah[...] = A
            l0 = h.launches
            F = q.ql_(Ah, handle=h)
            assert h.launches > l0
            outs.append((F.factors.copy(), F.tau.copy()))
        Bq = np.asfortranarray(rng.standard_normal((n, 4)))
        QtB = q.lmul_(F.Q.T, Bq.copy(order="F"))                            # T cache refilled by the replay (validated by content)
        Fp = q.ql(np.asfortranarray(A1), handle=h)                          # pageable copy: plain launches, same schedule
    finally:
        for name, v in (("host_split_min", 8192), ("la_min", 4096), ("chunk_cols", 1024)):
            h.set_option(name, v)
    for k in (1, 2, 4):
        assert np.array_equal(outs[0][0], outs[k][0]) and np.array_equal(outs[0][1], outs[k][1])
    assert np.array_equal(outs[0][0], Fp.factors) and np.array_equal(outs[0][1], Fp.tau)
    lapack.set_num_threads(16)
    for A, (Fa, ta) in ((A1, outs[0]), (A2, outs[3])):
        Fo, tauo = lapack.geqlf(np.asfortranarray(A.copy()))
        assert np.abs(ta - tauo).max() <= 50 * n * EPS
        assert np.abs(Fa - Fo).max() <= 50 * n * EPS * np.linalg.norm(A)
    ref = lapack.ormql("L", "T", outs[4][0], outs[4][1], Bq.copy(order="F"))
    assert np.abs(QtB - ref).max() <= 50 * n * EPS * np.linalg.norm(Bq)


def test_host_register_pins_a_numpy_matrix(q, h):
    """qlb200_host_register / _unregister (Handle.pin / unpin, QLB200.pin!): a pageable NumPy matrix page-locked in place
    uploads at PCIe speed and takes the graph replay of the large host-pointer ql!; same bits as the pageable call."""
    n = 2304
    A0 = np.asfortranarray(np.random.default_rng(11).standard_normal((n, n)))
    for name, v in (("host_split_min", 1024), ("la_min", 1024), ("chunk_cols", 512)):
        h.set_option(name, v)
    try:
        Fp = q.ql(A0, handle=h)                         # pageable
        rate_pageable = h.upload_rate()
        A = A0.copy(order="F")
        h.pin(A)
        h.pin(A)                                        # registering twice is not an error
        try:
            outs = []
            for _ in range(3):
                A[...] = A0
                F = q.ql_(A, handle=h)
                outs.append((F.factors.copy(), F.tau.copy()))
        finally:
            h.unpin(A)
            h.unpin(A)                                  # neither is releasing twice
        A[...] = A0
        Fu = q.ql_(A, handle=h)                         # pageable again: plain launches, the cached graph is gone
    finally:
        for name, v in (("host_split_min", 8192), ("la_min", 4096), ("chunk_cols", 1024)):
            h.set_option(name, v)
    assert rate_pageable > 0.1
    for f, t in outs:
        assert np.array_equal(f, Fp.factors) and np.array_equal(t, Fp.tau)
    assert np.array_equal(Fu.factors, Fp.factors) and np.array_equal(Fu.tau, Fp.tau)
    with pytest.raises(Exception):
        h.call("qlb200_host_register", 0, 8)


def _scaled_cmp(F, tau, Fo, tauo, A, k_rows_are_reflectors=False):
    """Elementwise comparison that stays meaningful when columns of A live at very different scales: every entry of the
    packed factors is compared relative to max(1, ...) of ITS column -- reflector entries are O(1), L entries carry the
    column's own scale -- at the usual 50*n*eps."""
    m, n = A.shape
    assert np.isfinite(F).all() and np.isfinite(tau).all()
    scale = np.maximum(np.abs(Fo).max(axis=1 if k_rows_are_reflectors else 0, keepdims=True), np.finfo(np.float64).tiny)
    scale = np.maximum(scale, np.abs(F).max(axis=1 if k_rows_are_reflectors else 0, keepdims=True))
    err = np.abs(F - Fo) / scale
    assert err.max() <= 50 * max(m, n) * EPS, err.max()
    assert np.abs(tau - tauo).max() <= 50 * max(m, n) * EPS


@pytest.mark.parametrize("shape", [(40, 40), (300, 200), (200, 300), (700, 700)])
@pytest.mark.parametrize("case", ["tiny", "huge", "tiny_col", "huge_col", "denormal_col", "mixed"])
def test_geqlf_extreme_scales_follow_lapack_larfg(q, h, shape, case):
    """LAPACK ?larfg (what src/ql.jl:72 runs) takes the norm in scaled form and rescales when |beta| < safmin: a matrix
    scaled by 1e-170 / 1e170, or with single columns at 1e-200 / 1e200 / in the denormal range, must factor like LAPACK
    (the blocked path's strip kernel falls to its scaled slow step whenever plain squares would leave the range)."""
    m, n = shape
    rng = np.random.default_rng(m * 13 + n + len(case))
    A = np.asfortranarray(rng.standard_normal(shape))
    if case == "tiny":
        A *= 1e-170
    elif case == "huge":
        A *= 1e170
    elif case == "tiny_col":
        A[:, n // 3] *= 1e-200
        A[:, n - 1] *= 1e-180
    elif case == "huge_col":
        A[:, n // 2] *= 1e200
    elif case == "denormal_col":
        A[:, n - 2] *= 1e-300
        A[:, 1] *= 1e-305
    else:
        A[:, ::3] *= 1e-165
        A[:, 1::3] *= 1e160
    F = q.ql(A, handle=h)
    Fo, tauo = lapack.geqlf(A)
    _scaled_cmp(F.factors, F.tau, Fo, tauo, A)


@pytest.mark.parametrize("case", ["tiny", "huge", "tiny_col"])
def test_gerqf_geqrf_extreme_scales(q, h, case):
    rng = np.random.default_rng(len(case))
    A = np.asfortranarray(rng.standard_normal((150, 260)))
    if case == "tiny":
        A *= 1e-170
    elif case == "huge":
        A *= 1e170
    else:
        A[40, :] *= 1e-200          # a tiny ROW is what a tiny column is to QL for RQ
    F, tau = q.rq_(A.copy(order="F"), handle=h)
    Fo, tauo = lapack.gerqf(A)
    _scaled_cmp(F, tau, Fo, tauo, A, k_rows_are_reflectors=True)
    At = np.asfortranarray(A.T)
    G, gt = q.qr_(At.copy(order="F"), handle=h)
    Go, gto = lapack.geqrf(At)
    _scaled_cmp(G, gt, Go, gto, At)


def test_geqlf_tall_path_extreme_scales(q, h):
    """m > 16384 (strips in global memory): the same ?larfg semantics."""
    rng = np.random.default_rng(4)
    A = np.asfortranarray(rng.standard_normal((17000, 24)))
    A[:, 5] *= 1e-200
    A[:, 20] *= 1e190
    A[:, 11] = 0.0
    F = q.ql(A, handle=h)
    Fo, tauo = lapack.geqlf(A)
    _scaled_cmp(F.factors, F.tau, Fo, tauo, A)
    for s in (1e-170, 1e170):
        B = np.asfortranarray(rng.standard_normal((16500, 8)) * s)
        F = q.ql(B, handle=h)
        Fo, tauo = lapack.geqlf(B)
        _scaled_cmp(F.factors, F.tau, Fo, tauo, B)


def test_public_entry_points_copy_column_major_device_input(q, h):
    """reference src/ql.jl:179-186: ql(A) = ql!(copy(A)); F \\ b works on a copy of b.  A column-major CUDA tensor is
    exactly the input whose `x.T.contiguous().T` is an alias of itself."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(11)
    n = 96
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T        # column-major
    b = torch.randn((3, n), dtype=torch.float64, device="cuda", generator=g).T
    A0, b0 = A.clone(), b.clone()
    F = q.ql(A, handle=h)
    torch.cuda.synchronize()
    assert F.factors.data_ptr() != A.data_ptr() and torch.equal(A, A0)
    x = F.solve(b)
    torch.cuda.synchronize()
    assert x.data_ptr() != b.data_ptr() and torch.equal(b, b0)
    assert (A0 @ x - b0).abs().max().item() <= 5000 * EPS * torch.linalg.cond(A0).item()
    Fu = q.ul(A, handle=h)
    torch.cuda.synchronize()
    assert torch.equal(A, A0) and Fu.factors.data_ptr() != A.data_ptr()
    xu = Fu.solve(b)
    torch.cuda.synchronize()
    assert torch.equal(b, b0) and (A0 @ xu - b0).abs().max().item() <= 5000 * EPS * torch.linalg.cond(A0).item()


def test_t_cache_is_validated_by_content(q, h):
    """The default T cache must never serve factors it does not belong to: (a) device factors edited IN PLACE in an entry
    no sampled fingerprint would see, tau unchanged; (b) other factors written into the SAME buffers; (c) host factors
    edited in place.  Each time lmul! must equal the answer computed with the cache switched off."""
    import torch

    n = 700
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T
    A2 = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T
    B = torch.randn((5, n), dtype=torch.float64, device="cuda", generator=g).T

    def apply(F, cache):
        h.set_option("tcache", cache)
        try:
            out = q.lmul_(F.Q.T, (B.T.clone()).T)
            torch.cuda.synchronize()
        finally:
            h.set_option("tcache", 1)
        return out

    buf = (A.T.clone()).T
    F = q.ql_(buf, handle=h)
    torch.cuda.synchronize()
    hit = apply(F, 1)
    assert torch.equal(hit, apply(F, 0))
    F = q.ql_(buf.copy_(A), handle=h)            # refactor (the off-run above dropped the cache)
    F.factors[333, 555] += 0.25                  # (a) one reflector entry, tau untouched
    torch.cuda.synchronize()
    got = apply(F, 1)
    assert torch.equal(got, apply(F, 0)) and not torch.equal(got, hit)
    F = q.ql_(buf.copy_(A), handle=h)
    h2 = q.Handle(0)
    other = q.ql_((A2.T.clone()).T, handle=h2)               # factors of another matrix, made elsewhere ...
    torch.cuda.synchronize()
    h2.close()
    F.factors.copy_(other.factors)                           # (b) ... land in the buffers the cache was keyed on
    F.tau.copy_(other.tau)
    torch.cuda.synchronize()
    assert torch.equal(apply(F, 1), apply(F, 0))
    # (c) host pointers
    rng = np.random.default_rng(8)
    Ah = np.asfortranarray(rng.standard_normal((n, n)))
    Bh = np.asfortranarray(rng.standard_normal((n, 2)))
    Fh = q.ql(Ah, handle=h)
    Fh.factors[100, 400] -= 0.5
    got = q.lmul_(Fh.Q.T, Bh.copy(order="F"))
    ref = cport.lmul(Fh.factors, Fh.tau, Bh, adjoint=True)
    assert np.abs(got - ref).max() <= 50 * n * EPS * np.abs(Bh).max() * np.sqrt(n)


@pytest.mark.parametrize("n,la,pn", [(2304, 1, 8192), (700, 1, 8192), (4608, 2, 8192), (2304, 1, 1), (4608, 2, 2048), (1100, 1, 512)])
def test_lookahead_ql_plain_captured_and_replayed_agree(q, h, n, la, pn):
    """The look-ahead ql! (panel j+1 factored on the panel SM partition while the GEMM partition applies panel j) called three
    times on the same buffers: plain launches, graph capture, graph replay -- bitwise the same factors, LAPACK parity, and
    the same bits as a fresh handle (the schedule is fixed: no run-to-run variation).  pn = "la_pnarrow_min": while the next
    panel starts at or right of that column the PANEL partition applies the finished panel to the next panel's columns itself
    (1: for every panel; 8192: never at these sizes; in between: the hand-over inside one factorization)."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(n)
    A = torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g).T
    buf = torch.empty((n, n), dtype=torch.float64, device="cuda").T
    outs = []
    h.set_option("lookahead", la)
    h.set_option("la_pnarrow_min", pn)
    try:
        for _ in range(3):
            buf.T.copy_(A.T)
            F = q.ql_(buf, handle=h)
            torch.cuda.synchronize()
            outs.append((F.factors.clone(), F.tau.clone()))
    finally:
        h.set_option("lookahead", 2)
        h.set_option("la_pnarrow_min", 8192)
    for f, t in outs[1:]:
        assert torch.equal(f, outs[0][0]) and torch.equal(t, outs[0][1])
    Ah = np.asfortranarray(A.cpu().numpy())
    lapack.set_num_threads(16)
    Fo, tauo = lapack.geqlf(Ah)
    tol = _tol((n, n), Ah)
    assert np.abs(outs[0][0].cpu().numpy() - Fo).max() <= tol and np.abs(outs[0][1].cpu().numpy() - tauo).max() <= tol
    # the T cache filled by the look-ahead path serves lmul!
    B = torch.randn((3, n), dtype=torch.float64, device="cuda", generator=g).T
    got = q.lmul_(F.Q.T, (B.T.clone()).T)
    torch.cuda.synchronize()
    ref = lapack.ormql("L", "T", np.asfortranarray(Fo), tauo, np.asfortranarray(B.cpu().numpy()))
    assert np.abs(got.cpu().numpy() - ref).max() <= 50 * n * EPS * np.abs(ref).max() * np.sqrt(n)
