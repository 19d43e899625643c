"""CPU check of the NumPy model of the device driver (tools/proto_blocked.py): the panel/strip
decomposition, Gram-based T and the larfb GEMM formulation that csrc/blocked.cu runs on the GPU
must reproduce LAPACK ?geqlf / ?ormql (the routines behind the reference's ql!, src/ql.jl:72)."""
import numpy as np
import pytest

from oracle import cport, lapack
from tools.proto_blocked import geqlf_blocked, ormql_blocked


@pytest.mark.parametrize("shape", [(10, 10), (10, 12), (12, 10), (3, 5), (100, 103), (200, 150), (300, 300), (150, 420), (1, 1)])
@pytest.mark.parametrize("nb,sb", [(128, 16), (32, 16), (48, 16)])
def test_proto_geqlf_matches_lapack(shape, nb, sb):
    A = np.random.default_rng(3).standard_normal(shape) / 2
    F, tau = geqlf_blocked(A, nb=nb, sb=sb)
    Fo, tauo = lapack.geqlf(A)
    t = 50 * max(shape) * np.finfo(float).eps * np.linalg.norm(A)
    assert np.abs(F - Fo).max() <= t and np.abs(tau - tauo).max() <= t


@pytest.mark.parametrize("shape", [(60, 60), (70, 40), (40, 70)])
@pytest.mark.parametrize("side", ["L", "R"])
@pytest.mark.parametrize("trans", ["N", "T"])
def test_proto_ormql_matches_lapack(shape, side, trans):
    rng = np.random.default_rng(4)
    A = rng.standard_normal(shape)
    F, tau = lapack.geqlf(A)
    m, n = shape
    k = min(m, n)
    C = rng.standard_normal((m, 7)) if side == "L" else rng.standard_normal((7, m))
    got = ormql_blocked(side, trans, F, tau, C, nb=32)
    ref = lapack.ormql(side, trans, np.asfortranarray(F[:, n - k:]), tau, C)
    assert np.abs(got - ref).max() <= 1e-12
    # and the reference's scalar loops (src/ql.jl:321-435)
    ref2 = cport.lmul(F, tau, C, adjoint=(trans == "T")) if side == "L" else cport.rmul(C, F, tau, adjoint=(trans == "T"))
    assert np.abs(got - ref2).max() <= 1e-12


@pytest.mark.parametrize("n,pad", [(8, 0), (16, 0), (32, 0), (48, 0), (64, 0), (32, 8), (32, 5), (40, 7), (64, 17), (64, 31)])
def test_proto_batched_wy_matches_lapack(n, pad):
    """NumPy model of the shared-memory batched kernel (csrc/batched_wy.cuh): rotating panel steps, compact-WY factor
    S from the Gram rows, C <- C - V S (V'C), zero-padded tile for orders that are not a multiple of 8."""
    from tools.proto_wy_batched import ql_wy

    nr = n - pad
    Ar = np.random.default_rng(n * 100 + pad).standard_normal((nr, nr))
    A = np.zeros((n, n))
    A[pad:, pad:] = Ar
    F, tau = ql_wy(A, pad)
    Fo, tauo = lapack.geqlf(np.asfortranarray(Ar))
    t = 50 * n * np.finfo(float).eps * np.linalg.norm(Ar)
    assert np.abs(F[pad:, pad:] - Fo).max() <= t and np.abs(tau[pad:] - tauo).max() <= t


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
@pytest.mark.parametrize("shape", [(9, 9), (12, 7), (7, 12)])
def test_lapack_binding_complex(dt, shape):
    """The complex half of the LAPACK binding (z/c geqlf, unmql: what src/ql.jl:72 calls for Complex BlasFloat) -- used as the
    parity reference of tests/test_eltypes_gpu.py: Q unitary, Q [0; L] = A, Q^H via 'C'."""
    m, n = shape
    rng = np.random.default_rng(m * 10 + n)
    A = np.asfortranarray((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt))
    F, tau = lapack.geqlf(A)
    k = min(m, n)
    V = np.asfortranarray(F[:, n - k:])
    Q = lapack.ormql("L", "N", V, tau, np.eye(m, dtype=dt))
    QH = lapack.ormql("L", "C", V, tau, np.eye(m, dtype=dt))
    eps = np.finfo(dt).eps
    assert np.abs(Q.conj().T @ Q - np.eye(m)).max() <= 50 * m * eps
    assert np.abs(QH - Q.conj().T).max() <= 50 * m * eps
    L = np.zeros((m, n), dtype=dt)
    L[m - k:, :] = np.tril(F[m - k:, :], max(n - m, 0))
    assert np.abs(Q @ L - A).max() <= 50 * max(m, n) * eps * np.linalg.norm(A)
    if k > 1:      # zlarfg: beta is real
        d = np.array([F[m - k + i, n - k + i] for i in range(1, k)])
        assert np.abs(d.imag).max() == 0


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
@pytest.mark.parametrize("shape", [(9, 9), (12, 7), (7, 12)])
def test_generic_loops_restatement_vs_lapack(dt, shape):
    """oracle/loops.py (the reference's generic lmul!/rmul! loops for any element type, src/ql.jl:318-435, incl. the conj
    placements) against LAPACK ?ormql / ?unmql on LAPACK-computed factors, and against the real C port."""
    from oracle import loops

    m, n = shape
    rng = np.random.default_rng(m + 31 * n)
    A = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if dt == np.complex128 else 0)
    A = np.asfortranarray(A.astype(dt))
    F, tau = lapack.geqlf(A)
    k = min(m, n)
    V = np.asfortranarray(F[:, n - k:])
    B = np.asfortranarray((rng.standard_normal((m, 3)) + (1j * rng.standard_normal((m, 3)) if dt == np.complex128 else 0)).astype(dt))
    C = np.asfortranarray((rng.standard_normal((3, m)) + (1j * rng.standard_normal((3, m)) if dt == np.complex128 else 0)).astype(dt))
    adj = "C" if dt == np.complex128 else "T"
    assert np.abs(loops.lmul(F, tau, B) - lapack.ormql("L", "N", V, tau, B)).max() <= 1e-12
    assert np.abs(loops.lmul(F, tau, B, adjoint=True) - lapack.ormql("L", adj, V, tau, B)).max() <= 1e-12
    assert np.abs(loops.rmul(C, F, tau) - lapack.ormql("R", "N", V, tau, C)).max() <= 1e-12
    assert np.abs(loops.rmul(C, F, tau, adjoint=True) - lapack.ormql("R", adj, V, tau, C)).max() <= 1e-12
    if dt == np.float64:
        assert np.abs(loops.lmul(F, tau, B) - cport.lmul(F, tau, B)).max() <= 1e-12
        assert np.abs(loops.rmul(C, F, tau, adjoint=True) - cport.rmul(C, F, tau, adjoint=True)).max() <= 1e-12
    if m == n:
        x = loops.ldiv_square(F, tau, B)
        assert np.abs(A @ x - B).max() <= 1e-10 * np.linalg.cond(A)
