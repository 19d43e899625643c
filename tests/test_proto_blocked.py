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
