"""ctypes front-end of the CPU oracle (oracle/liboracle.so, built from ql_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The product package never imports this module.

All functions take/return NumPy arrays in column-major (Fortran) order, mirroring the reference's
Julia `Matrix` layout; reference file:line citations live next to each C function in
oracle/ql_oracle_impl.h.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with gcc (make).  Returns the path."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ql_oracle.c", "ql_oracle_impl.h", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_dulfact.restype = _i64
        _LIB.orc_sulfact.restype = _i64
        _LIB.orc_djl_reflector.restype = ctypes.c_double
        _LIB.orc_sjl_reflector.restype = ctypes.c_float
    return _LIB


def _pfx(a: np.ndarray) -> str:
    if a.dtype == np.float64:
        return "orc_d"
    if a.dtype == np.float32:
        return "orc_s"
    raise TypeError(f"oracle supports float32/float64, got {a.dtype}")


def _f(a, dtype=None) -> np.ndarray:
    a = np.array(a, dtype=dtype, order="F", copy=True)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64, order="F")
    return a


def _ptr(a: np.ndarray):
    return _vp(a.ctypes.data)


def _ld(a: np.ndarray) -> int:
    return max(1, a.strides[1] // a.itemsize) if a.ndim == 2 else max(1, a.shape[0])


def larfg(alpha: float, x, dtype=np.float64):
    """LAPACK ?larfg on (alpha, x).  Returns (beta, v, tau)."""
    x = _f(np.atleast_1d(x), dtype)
    al = np.array([alpha], dtype=x.dtype)
    tau = np.zeros(1, dtype=x.dtype)
    getattr(lib(), _pfx(x) + "larfg")(_i64(x.size + 1), _ptr(al), _ptr(x), _i64(1), _ptr(tau))
    return al[0], x, tau[0]


def jl_reflector(x, dtype=np.float64):
    """Julia reflector!(x): returns (x_overwritten, tau)."""
    x = _f(np.atleast_1d(x), dtype)
    tau = getattr(lib(), _pfx(x) + "jl_reflector")(_i64(x.size), _ptr(x), _i64(1))
    return x, x.dtype.type(tau)


def geql2(A, convention: int = 0):
    """Unblocked QL (src/ql.jl:56-68).  Returns (factors, tau).  convention 0 = LAPACK larfg
    (the Float64/Float32 production path, src/ql.jl:72), 1 = Julia generic reflector!."""
    A = _f(A)
    m, n = A.shape
    tau = np.zeros(min(m, n), dtype=A.dtype)
    getattr(lib(), _pfx(A) + "geql2")(_i64(m), _i64(n), _ptr(A), _i64(_ld(A)), _ptr(tau), _int(convention))
    return A, tau


def geql2_batched(A, convention: int = 0):
    """A: (batch, n, m) C-order array == batch column-major m x n matrices (Julia Array{T,3}
    (m,n,batch) memory order).  Returns (factors, tau[batch, min(m,n)])."""
    A = np.array(A, copy=True, order="C")
    batch, n, m = A.shape
    tau = np.zeros((batch, min(m, n)), dtype=A.dtype)
    getattr(lib(), _pfx(A) + "geql2_batched")(_i64(m), _i64(n), _ptr(A), _i64(m), _i64(m * n), _ptr(tau),
                                              _i64(min(m, n)), _i64(batch), _int(convention))
    return A, tau


def getL(factors):
    """F.L (src/ql.jl:214-217)."""
    A = np.asfortranarray(factors)
    m, n = A.shape
    L = np.zeros((min(m, n), n), dtype=A.dtype, order="F")
    getattr(lib(), _pfx(A) + "getL")(_i64(m), _i64(n), _ptr(A), _i64(_ld(A)), _ptr(L), _i64(_ld(L)))
    return L


class DimensionMismatch(ValueError):
    pass


def _apply(name, factors, tau, B, adjoint):
    A = np.asfortranarray(factors)
    tau = np.ascontiguousarray(tau, dtype=A.dtype)
    vec = np.ndim(B) == 1
    Bm = _f(np.reshape(B, (-1, 1)) if vec else B, A.dtype)
    rc = getattr(lib(), _pfx(A) + name)(_i64(A.shape[0]), _i64(A.shape[1]), _ptr(A), _i64(_ld(A)), _ptr(tau),
                                         _i64(Bm.shape[0]), _i64(Bm.shape[1]), _ptr(Bm), _i64(_ld(Bm)),
                                         _int(1 if adjoint else 0))
    if rc != 0:
        raise DimensionMismatch(f"{name}: factors {A.shape} vs operand {Bm.shape}")
    return Bm[:, 0].copy() if vec else Bm


def lmul(factors, tau, B, adjoint=False):
    """lmul!(Q,B) / lmul!(Q',B) scalar loops (src/ql.jl:321-347 / 350-377)."""
    return _apply("lmul_q", factors, tau, B, adjoint)


def rmul(C, factors, tau, adjoint=False):
    """rmul!(C,Q) / rmul!(C,Q') (src/ql.jl:381-406 / 409-435)."""
    return _apply("rmul_q", factors, tau, C, adjoint)


def matrix_q(factors, tau):
    """Matrix(Q) = lmul!(Q, I(m, min(m,n))) (src/ql.jl:264-266)."""
    A = np.asfortranarray(factors)
    m, n = A.shape
    Q = np.zeros((m, min(m, n)), dtype=A.dtype, order="F")
    tau = np.ascontiguousarray(tau, dtype=A.dtype)
    getattr(lib(), _pfx(A) + "matrix_q")(_i64(m), _i64(n), _ptr(A), _i64(_ld(A)), _ptr(tau), _ptr(Q), _i64(_ld(Q)))
    return Q


def square_q(factors, tau):
    """squareQ helper of the reference tests (test/test_qr.jl:19): lmul!(Q, I(m,m))."""
    m = np.shape(factors)[0]
    return lmul(factors, tau, np.eye(m, dtype=np.asarray(factors).dtype), adjoint=False)


def ldiv(factors, tau, B):
    """ldiv!(F::QL,B), square only (src/ql.jl:440-447,467)."""
    A = np.asfortranarray(factors)
    n = A.shape[0]
    if A.shape[0] != A.shape[1]:
        raise DimensionMismatch("oracle ldiv: the reference pins only the square solve (test/test_ql.jl:69)")
    vec = np.ndim(B) == 1
    Bm = _f(np.reshape(B, (-1, 1)) if vec else B, A.dtype)
    if Bm.shape[0] != n:
        raise DimensionMismatch(f"ldiv: factors {A.shape} vs rhs {Bm.shape}")
    tau = np.ascontiguousarray(tau, dtype=A.dtype)
    rc = getattr(lib(), _pfx(A) + "ldiv_ql")(_i64(n), _ptr(A), _i64(_ld(A)), _ptr(tau), _i64(Bm.shape[1]),
                                             _ptr(Bm), _i64(_ld(Bm)))
    if rc > 0:
        raise ZeroDivisionError(f"singular L at {rc}")
    return Bm[:, 0].copy() if vec else Bm


def gerq2(A, convention: int = 0):
    """Unblocked RQ (src/rq.jl:353-380 / LAPACK ?gerq2).  Returns (factors, tau)."""
    A = _f(A)
    m, n = A.shape
    tau = np.zeros(min(m, n), dtype=A.dtype)
    getattr(lib(), _pfx(A) + "gerq2")(_i64(m), _i64(n), _ptr(A), _i64(_ld(A)), _ptr(tau), _int(convention))
    return A, tau


def rq_lmul(factors, tau, B, adjoint=False):
    """lmul!(Q::RQPackedQ,B) / lmul!(Q',B) generic loops (src/rq.jl:138-167 / 203-233)."""
    return _apply("rq_lmul", factors, tau, B, adjoint)


def rq_rmul(C, factors, tau, adjoint=False):
    """rmul!(C,Q::RQPackedQ) / rmul!(C,Q') generic loops (src/rq.jl:243-271 / 292-321)."""
    return _apply("rq_rmul", factors, tau, C, adjoint)


def ulfact(A, pivot: bool = True):
    """generic_ulfact! (src/ul.jl:148-196).  Returns (factors, ipiv (1-based int64), info)."""
    A = _f(A)
    m, n = A.shape
    ipiv = np.zeros(min(m, n), dtype=np.int64)
    info = getattr(lib(), _pfx(A) + "ulfact")(_i64(m), _i64(n), _ptr(A), _i64(_ld(A)), _ptr(ipiv), _int(int(pivot)))
    return A, ipiv, int(info)


def ul_ldiv(factors, ipiv, B):
    """ldiv!(F::UL,B) (src/ul.jl:390-393), square."""
    A = np.asfortranarray(factors)
    n = A.shape[0]
    vec = np.ndim(B) == 1
    Bm = _f(np.reshape(B, (-1, 1)) if vec else B, A.dtype)
    ipiv = np.ascontiguousarray(ipiv, dtype=np.int64)
    rc = getattr(lib(), _pfx(A) + "ul_ldiv")(_i64(n), _ptr(A), _i64(_ld(A)), _ptr(ipiv), _i64(Bm.shape[1]),
                                             _ptr(Bm), _i64(_ld(Bm)))
    if rc > 0:
        raise ZeroDivisionError(f"singular L at {rc}")
    return Bm[:, 0].copy() if vec else Bm
