"""ctypes binding of the LAPACK routines the reference itself calls, taken from the OpenBLAS that
ships inside SciPy's wheel (scipy.libs/libscipy_openblas*.so, symbols `scipy_<name>_`, LP64).

TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Never imported by the product package.

Why this exists: for Float64/Float32 the reference's factorization arithmetic is not in its own
tree -- `ql!` bottoms out in `LAPACK.geqlf!` (src/ql.jl:72), `rq!` in `LAPACK.gerqf!`
(src/rq.jl:401) and RQ's Q-apply in `LAPACK.ormrq!` (src/rq.jl:130-137).  Julia is not installed
in this image, but the *same* reference-LAPACK Fortran routines are exported by SciPy's OpenBLAS
(0.3.31.dev here; Julia >= 1.10 bundles a 0.3.2x build -- identical LAPACK sources, results differ
only by BLAS-kernel rounding).  They pin the C restatement (tests/test_oracle.py), generate the
golden fixtures (tests/golden/make_golden.py) and serve as the timed CPU baseline (bench.py).
"""
from __future__ import annotations

import ctypes
import glob
import os
from ctypes import byref, c_char, c_double, c_float, c_int, c_void_p

import numpy as np

_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        import scipy

        pats = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so")
        cands = sorted(p for p in glob.glob(pats) if "64_" not in os.path.basename(p))
        if not cands:
            raise OSError("scipy-openblas shared library not found (needed as the LAPACK oracle)")
        _LIB = ctypes.CDLL(cands[0])
        _LIB.scipy_openblas_get_num_threads.restype = c_int
        _LIB.scipy_openblas_get_config.restype = ctypes.c_char_p
    return _LIB


def set_num_threads(n: int) -> None:
    lib().scipy_openblas_set_num_threads(c_int(int(n)))


def get_num_threads() -> int:
    return int(lib().scipy_openblas_get_num_threads())


def config() -> str:
    return lib().scipy_openblas_get_config().decode()


def _p(a):
    return c_void_p(a.ctypes.data)


def _pre(a):
    return {np.dtype(np.float64): "d", np.dtype(np.float32): "s", np.dtype(np.complex128): "z", np.dtype(np.complex64): "c"}[a.dtype]


def _real(a):
    return c_double if a.dtype == np.float64 else c_float


def _fact(name, A, inplace=False):
    """Common driver for ?geqlf / ?gerqf / ?geqrf: two-call workspace query like Julia's wrapper."""
    A = A if inplace else np.array(A, order="F", copy=True)
    assert A.flags.f_contiguous or A.strides[0] == A.itemsize
    m, n = A.shape
    lda = max(1, A.strides[1] // A.itemsize)
    tau = np.zeros(min(m, n), dtype=A.dtype)
    fn = getattr(lib(), f"scipy_{_pre(A)}{name}_")
    info = c_int(0)
    wq = np.zeros(1, dtype=A.dtype)
    fn(byref(c_int(m)), byref(c_int(n)), _p(A), byref(c_int(lda)), _p(tau), _p(wq), byref(c_int(-1)), byref(info))
    lwork = max(1, int(np.real(wq[0])))
    work = np.empty(lwork, dtype=A.dtype)
    fn(byref(c_int(m)), byref(c_int(n)), _p(A), byref(c_int(lda)), _p(tau), _p(work), byref(c_int(lwork)), byref(info))
    if info.value != 0:
        raise ValueError(f"{name}: info={info.value}")
    return A, tau


def geqlf(A, inplace=False):
    """LAPACK ?geqlf -- the routine behind `ql!` for BlasFloat strided matrices (src/ql.jl:72)."""
    return _fact("geqlf", A, inplace)


def gerqf(A, inplace=False):
    """LAPACK ?gerqf -- the routine behind `rq!` (src/rq.jl:401)."""
    return _fact("gerqf", A, inplace)


def geqrf(A, inplace=False):
    """LAPACK ?geqrf (used for the flip identity QL <-> QR, test/test_ql.jl:7-13)."""
    return _fact("geqrf", A, inplace)


def _orm(name, side, trans, A, tau, C):
    A = np.asfortranarray(A)
    C = np.array(C, order="F", copy=True, dtype=A.dtype)
    vec = C.ndim == 1
    if vec:
        C = C.reshape(-1, 1, order="F")
    m, n = C.shape
    k = tau.size
    lda = max(1, A.strides[1] // A.itemsize)
    ldc = max(1, C.strides[1] // C.itemsize)
    tau = np.ascontiguousarray(tau, dtype=A.dtype)
    if _pre(A) in "zc":
        name = name.replace("orm", "unm")          # ?unmql / ?unmqr / ?unmrq (trans 'N' or 'C')
    fn = getattr(lib(), f"scipy_{_pre(A)}{name}_")
    info = c_int(0)
    wq = np.zeros(1, dtype=A.dtype)
    s, t = c_char(side.encode()), c_char(trans.encode())
    args = lambda w, lw: (byref(s), byref(t), byref(c_int(m)), byref(c_int(n)), byref(c_int(k)), _p(A),
                          byref(c_int(lda)), _p(tau), _p(C), byref(c_int(ldc)), _p(w), byref(c_int(lw)), byref(info),
                          ctypes.c_size_t(1), ctypes.c_size_t(1))
    fn(*args(wq, -1))
    lwork = max(1, int(np.real(wq[0])))
    work = np.empty(lwork, dtype=A.dtype)
    fn(*args(work, lwork))
    if info.value != 0:
        raise ValueError(f"{name}: info={info.value}")
    return C[:, 0].copy() if vec else C


def ormql(side, trans, A, tau, C):
    """LAPACK ?ormql ("what LAPACK would do" line next to the reference's scalar lmul! loops)."""
    return _orm("ormql", side, trans, A, tau, C)


def ormqr(side, trans, A, tau, C):
    """LAPACK ?ormqr -- QRPackedQ's lmul!/rmul! for BlasFloat (LinearAlgebra stdlib; QR of src/qr.jl:72)."""
    return _orm("ormqr", side, trans, A, tau, C)


def ormrq(side, trans, A, tau, C):
    """LAPACK ?ormrq -- RQ's Q-apply in the reference (src/rq.jl:130-137,183-202)."""
    return _orm("ormrq", side, trans, A, tau, C)


def getrf(A):
    """LAPACK ?getrf (UL == flip o getrf o flip).  Returns (LU, ipiv 1-based, info)."""
    A = np.array(A, order="F", copy=True)
    m, n = A.shape
    lda = max(1, A.strides[1] // A.itemsize)
    ipiv = np.zeros(min(m, n), dtype=np.int32)
    info = c_int(0)
    getattr(lib(), f"scipy_{_pre(A)}getrf_")(byref(c_int(m)), byref(c_int(n)), _p(A), byref(c_int(lda)), _p(ipiv),
                                             byref(info))
    return A, ipiv.astype(np.int64), info.value


def geqlf_loop(A3, threads: int = 1):
    """The reference-equivalent batched workload `for i: ql!(view(A,:,:,i))`:
    A3 is (batch, n, m) C-order == batch column-major m x n matrices.  In place.  Returns tau."""
    batch, n, m = A3.shape
    tau = np.zeros((batch, min(m, n)), dtype=A3.dtype)
    fn = getattr(lib(), f"scipy_{_pre(A3)}geqlf_")
    lwork = 64 * max(m, n)
    work = np.empty(lwork, dtype=A3.dtype)
    info = c_int(0)
    cm, cn, clw = c_int(m), c_int(n), c_int(lwork)
    a0, t0 = A3.ctypes.data, tau.ctypes.data
    sa, st = A3.strides[0], tau.strides[0]
    wp = _p(work)
    for b in range(batch):
        fn(byref(cm), byref(cn), c_void_p(a0 + b * sa), byref(cm), c_void_p(t0 + b * st), wp, byref(clw), byref(info))
    return tau
