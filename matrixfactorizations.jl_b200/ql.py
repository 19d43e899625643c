"""Host-side mirror of MatrixFactorizations.jl's QL interface on top of libqlb200.so.

The reference's host language is Julia (absent from this image); its shim is in
julia/QLB200.jl and INTEGRATION.md.  This module is the executable host side: the same names,
argument meaning and error behaviour as reference src/ql.jl, each call one C-ABI entry point.

    reference (Julia)                 here
    ---------------------------------------------------------------
    ql!(A)            src/ql.jl:116   ql_(A)            in place, returns QL
    ql(A)             src/ql.jl:179   ql(A)             copies first (src/ql.jl:181-186)
    QL(factors, tau)  src/ql.jl:37    QL(factors, tau)
    F.L / F.Q         src/ql.jl:214   F.L / F.Q
    QLPackedQ         src/ql.jl:245   QLPackedQ  (Q.T is the adjoint, Matrix(Q) -> Q.matrix())
    lmul!(Q,B)/(Q',B) src/ql.jl:321   lmul_(Q, B) / lmul_(Q.T, B)
    rmul!(C,Q)/(C,Q') src/ql.jl:381   rmul_(C, Q) / rmul_(C, Q.T)
    ldiv!(F,B), F\\b   src/ql.jl:438   ldiv_(F, B), F.solve(b)
    rq!(A)            src/rq.jl:401   rq_(A)
    (new) batched     SURVEY 0.6      ql_batched_(A3), lmul_batched_, ldiv_batched_

Arrays are column-major like Julia's: NumPy arrays must be Fortran-ordered (host path: the library
stages H2D/D2H itself); torch CUDA tensors must have stride(0) == 1 (device path, `_dev` entry
points, asynchronous on the handle's stream -- call handle.sync()).  Batched arrays have shape
(batch, n_cols, n_rows) C-contiguous == Julia Array{T,3}(n_rows, n_cols, batch) memory.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch, Handle, SingularException, default_handle  # noqa: F401

_c = ctypes


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _pfx(x) -> str:
    name = str(x.dtype)
    if name.endswith("float64"):
        return "d"
    if name.endswith("float32"):
        return "s"
    if name.endswith("complex128"):
        return "z"
    if name.endswith("complex64"):
        return "c"
    raise TypeError(f"qlb200 supports the BlasFloat element types Float64/Float32/ComplexF64/ComplexF32 (the reference's LAPACK "
                    f"path, src/ql.jl:72), got {x.dtype}")


_ITEMSIZE = {"d": 8, "s": 4, "z": 16, "c": 8}


class _Mat:
    """Column-major view descriptor: pointer, rows, cols, leading dimension, host/device."""

    def __init__(self, x, what="A"):
        if _is_torch(x):
            if not x.is_cuda:
                raise TypeError(f"{what}: torch tensors must live on a CUDA device (use NumPy for host data)")
            if x.dim() == 1:
                self.m, self.n, self.ld = x.shape[0], 1, max(1, x.shape[0])
                if x.shape[0] > 1 and x.stride(0) != 1:
                    raise ValueError(f"{what}: vector must be contiguous")
            else:
                self.m, self.n = x.shape
                if self.m > 1 and x.stride(0) != 1:
                    raise ValueError(f"{what}: device matrices must be column-major (stride(0) == 1)")
                self.ld = max(1, x.stride(1)) if self.n > 1 else max(1, self.m)
            self.ptr, self.dev, self.device_index = x.data_ptr(), True, x.device.index or 0
        else:
            if not isinstance(x, np.ndarray):
                raise TypeError(f"{what}: expected numpy.ndarray or torch.Tensor")
            if x.ndim == 1:
                self.m, self.n, self.ld = x.shape[0], 1, max(1, x.shape[0])
                if x.shape[0] > 1 and x.strides[0] != x.itemsize:
                    raise ValueError(f"{what}: vector must be contiguous")
            else:
                self.m, self.n = x.shape
                if self.m > 1 and x.strides[0] != x.itemsize:
                    raise ValueError(f"{what}: host matrices must be column-major (Fortran order)")
                self.ld = max(1, x.strides[1] // x.itemsize) if self.n > 1 else max(1, self.m)
            if not x.flags.writeable and what != "A(const)":
                pass
            self.ptr, self.dev, self.device_index = x.ctypes.data, False, 0
        self.pfx = _pfx(x)
        self.itemsize = _ITEMSIZE[self.pfx]
        self.obj = x

    @property
    def sfx(self):
        return "_dev" if self.dev else ""


def _handle(handle, *mats):
    """The handle to use; for device operands its stream follows torch's current stream so the
    library's kernels are ordered with the caller's torch work."""
    dev, on_dev = 0, False
    for m in mats:
        if m.dev:
            dev, on_dev = m.device_index, True
    h = handle if handle is not None else default_handle(dev)
    if on_dev:
        import torch

        h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    return h


def _same_kind(*mats):
    if len({(m.dev, m.pfx) for m in mats}) != 1:
        raise TypeError("all operands must share element type and residency (all host or all device)")


def _new_like(x, n):
    if _is_torch(x):
        import torch

        return torch.empty(n, dtype=x.dtype, device=x.device)
    return np.empty(n, dtype=x.dtype)


def _copy_colmajor(x):
    """A fresh column-major copy (never an alias of x: `x.T.contiguous().T` returns x itself when x is already
    column-major, which would turn ql(A) / F.solve(b) into their in-place forms -- reference src/ql.jl:179-186 copies)."""
    if _is_torch(x):
        import torch

        if x.dim() != 2:
            return x.clone()
        out = torch.empty_strided(tuple(x.shape), (1, max(1, x.shape[0])), dtype=x.dtype, device=x.device)
        out.copy_(x)
        return out
    return np.array(x, order="F", copy=True)


class QLPackedQ:
    """The packed orthogonal factor (reference src/ql.jl:245-262): aliases factors and tau."""

    def __init__(self, factors, tau, adjoint=False, handle=None):
        self.factors, self.tau, self.adjoint, self.handle = factors, tau, adjoint, handle

    @property
    def T(self):
        return QLPackedQ(self.factors, self.tau, not self.adjoint, self.handle)

    adj = T

    @property
    def shape(self):            # src/ql.jl:270: size(Q) = (m, m)
        m = self.factors.shape[0]
        return (m, m)

    def matrix(self):
        """Matrix(Q) = lmul!(Q, Matrix(I, m, min(m,n))) (src/ql.jl:264-266)."""
        m, n = self.factors.shape
        k = min(m, n)
        if _is_torch(self.factors):
            import torch

            E = torch.zeros((k, m), dtype=self.factors.dtype, device=self.factors.device).T
            E.T.fill_diagonal_(1.0)
        else:
            E = np.zeros((m, k), dtype=self.factors.dtype, order="F")
            E[np.arange(k), np.arange(k)] = 1
        return lmul_(self, E)

    def __matmul__(self, B):
        """Q*B with the zero-pad-ON-TOP promotion of src/ql.jl:277-290."""
        m = self.factors.shape[0]
        k = min(self.factors.shape)
        vec = B.ndim == 1 if not _is_torch(B) else B.dim() == 1
        rows = B.shape[0]
        if rows not in (m, k):
            raise DimensionMismatch(f"first dimension of matrix must have size either {m} or {k}")
        if _is_torch(B):
            import torch

            cols = 1 if vec else B.shape[1]
            Bm = torch.zeros((cols, m), dtype=B.dtype, device=B.device).T
            Bm[m - rows:, :] = B.reshape(rows, cols)
        else:
            cols = 1 if vec else B.shape[1]
            Bm = np.zeros((m, cols), dtype=self.factors.dtype, order="F")
            Bm[m - rows:, :] = np.reshape(B, (rows, cols))
        out = lmul_(self, Bm)
        return out[:, 0] if vec else out


class QL:
    """QL factorization object (reference src/ql.jl:37-49): `factors` (m x n packed) + `tau`."""

    def __init__(self, factors, tau, handle=None):
        self.factors, self.tau, self.handle = factors, tau, handle

    @property
    def shape(self):
        return tuple(self.factors.shape)

    @property
    def Q(self):                 # getQ, src/ql.jl:218
        return QLPackedQ(self.factors, self.tau, False, self.handle)

    @property
    def L(self):                 # getL, src/ql.jl:214-217: bottom min(m,n) rows, tril(., max(n-m,0))
        m, n = self.factors.shape
        k = min(m, n)
        if _is_torch(self.factors):
            import torch

            return torch.tril(self.factors[m - k:, :], diagonal=max(n - m, 0))
        return np.tril(self.factors[m - k:, :], max(n - m, 0))

    def __iter__(self):          # destructuring Q, L = F (src/ql.jl:52-54)
        yield self.Q
        yield self.L

    def solve(self, b):
        r"""F \ b (src/ql.jl:490-502), square only -- the reference's rectangular branches are
        @test_broken (test/test_ql.jl:148)."""
        m, n = self.factors.shape
        if m < n:                                # wide: X = zeros(n, ...); X[1:m, :] = b (src/ql.jl:498-499); minimum-norm solution
            if b.shape[0] != m:
                raise DimensionMismatch(f"left hand side has {m} rows, but right hand side has {b.shape[0]} rows")
            if _is_torch(b):
                import torch

                cols = 1 if b.dim() == 1 else b.shape[1]
                X = torch.zeros((cols, n), dtype=b.dtype, device=b.device).T
                X[:m, :] = b.reshape(m, cols)
            else:
                cols = 1 if b.ndim == 1 else b.shape[1]
                X = np.zeros((n, cols), dtype=self.factors.dtype, order="F")
                X[:m, :] = np.reshape(b, (m, cols))
            ldiv_(self, X)
            return X[:, 0] if (b.dim() if _is_torch(b) else b.ndim) == 1 else X
        X = _copy_colmajor(b)
        ldiv_(self, X)
        return X if m == n else X[:n]            # `_cut_B(X, 1:n)`, src/ql.jl:501 (tall: the least-squares solution)


def ql_(A, handle=None) -> QL:
    """ql!(A) for a BlasFloat strided matrix: QL(LAPACK.geqlf!(A)...) (src/ql.jl:72,114-117)."""
    a = _Mat(A)
    if (A.ndim if not _is_torch(A) else A.dim()) != 2:
        raise ArgumentError("ql_ expects a matrix")
    h = _handle(handle, a)
    tau = _new_like(A, min(a.m, a.n))
    t = _Mat(tau, "tau")
    h.call(f"qlb200_{a.pfx}geqlf{a.sfx}", a.m, a.n, a.ptr, a.ld, t.ptr)
    return QL(A, tau, handle)


def ql(A, handle=None) -> QL:
    """ql(A): copy, then ql! (src/ql.jl:179-186).  Vectors/scalars become n x 1 (src/ql.jl:187-191)."""
    if not _is_torch(A):
        A = np.asarray(A)
        if A.dtype not in (np.float32, np.float64, np.complex64, np.complex128):
            A = A.astype(np.complex128 if np.iscomplexobj(A) else np.float64)        # _qleltype promotion (src/ql.jl:119)
        if A.ndim < 2:
            A = A.reshape(-1, 1)
    return ql_(_copy_colmajor(A), handle)


def rq_(A, handle=None):
    """rq!(A) = RQ(LAPACK.gerqf!(A)...) (src/rq.jl:401).  Returns (factors, tau)."""
    a = _Mat(A)
    h = _handle(handle, a)
    tau = _new_like(A, min(a.m, a.n))
    h.call(f"qlb200_{a.pfx}gerqf{a.sfx}", a.m, a.n, a.ptr, a.ld, _Mat(tau, "tau").ptr)
    return A, tau


def qr_(A, handle=None):
    """qrunblocked!(A) for a BlasFloat strided matrix = QR(LAPACK.geqrf!(A, tau)...) (src/qr.jl:72-76), computed by the
    QL kernels through the flip identity (test/test_ql.jl:7-13).  Returns (factors, tau)."""
    a = _Mat(A)
    if a.pfx != "d":
        raise TypeError("qr_: Float64 only")
    h = _handle(handle, a)
    tau = _new_like(A, min(a.m, a.n))
    h.call(f"qlb200_dgeqrf{a.sfx}", a.m, a.n, a.ptr, a.ld, _Mat(tau, "tau").ptr)
    return A, tau


def qr_solve(factors, tau, b, handle=None):
    r"""`QR(factors, tau) \ b` for m >= n (src/qr.jl:193-204): exact for square, least squares for tall.  Returns X (n x nrhs)."""
    X = _copy_colmajor(b)
    f, bm, t = _Mat(factors, "A(const)"), _Mat(X, "B"), _Mat(tau, "tau")
    _same_kind(f, bm, t)
    if f.m < f.n:
        raise DimensionMismatch("qr_solve: m >= n only")
    if bm.m != f.m:
        raise DimensionMismatch("Both inputs should have the same number of rows")
    h = _handle(handle, f)
    h.call(f"qlb200_dqrsolve{f.sfx}", f.m, f.n, bm.n, f.ptr, f.ld, t.ptr, bm.ptr, bm.ld)
    return X[: f.n]


class QRPackedQ:
    """QR's packed orthogonal factor: reflector i below the diagonal of column i of `factors` (m x n), size (m, m)."""

    def __init__(self, factors, tau, adjoint=False, handle=None):
        self.factors, self.tau, self.adjoint, self.handle = factors, tau, adjoint, handle

    @property
    def T(self):
        return QRPackedQ(self.factors, self.tau, not self.adjoint, self.handle)

    @property
    def shape(self):
        m = self.factors.shape[0]
        return (m, m)


def _ormqr(side, Q: QRPackedQ, C):
    f, c, t = _Mat(Q.factors, "A(const)"), _Mat(C, "C"), _Mat(Q.tau, "tau")
    _same_kind(f, c, t)
    h = _handle(Q.handle, f)
    mA, nA = f.m, f.n
    k = min(mA, nA)
    nq = c.m if side == "L" else c.n
    if nq != mA:
        raise DimensionMismatch(f"matrix A has dimensions ({mA},{nA}) but {'B' if side == 'L' else 'A'} has dimensions ({c.m}, {c.n})")
    h.call(f"qlb200_dormqr{f.sfx}", side.encode(), (b"T" if Q.adjoint else b"N"), c.m, c.n, k, f.ptr, f.ld, t.ptr, c.ptr, c.ld)
    return C


class RQPackedQ:
    """RQ's packed orthogonal factor (reference src/rq.jl:59-72): aliases factors (m x n) and tau; the k =
    min(m,n) reflectors are the LAST k rows of `factors`.  size(Q) = (n, n) (src/rq.jl:95)."""

    def __init__(self, factors, tau, adjoint=False, handle=None):
        self.factors, self.tau, self.adjoint, self.handle = factors, tau, adjoint, handle

    @property
    def T(self):
        return RQPackedQ(self.factors, self.tau, not self.adjoint, self.handle)

    @property
    def shape(self):
        n = self.factors.shape[1]
        return (n, n)


def _ormrq(side, Q: RQPackedQ, C):
    """lmul!/rmul! with RQ's Q: LAPACK.ormrq! (src/rq.jl:130-137,183-202,234-242,272-291)."""
    f, c, t = _Mat(Q.factors, "A(const)"), _Mat(C, "C"), _Mat(Q.tau, "tau")
    _same_kind(f, c, t)
    h = _handle(Q.handle, f)
    mA, nA = f.m, f.n
    k = min(mA, nA)
    nq = c.m if side == "L" else c.n
    if nq != nA:
        raise DimensionMismatch(f"matrix A has dimensions ({mA},{nA}) but {'B' if side == 'L' else 'A'} has dimensions ({c.m}, {c.n})")
    aptr = f.ptr + (mA - k) * f.itemsize  # first of the last k rows
    cplx = f.pfx in "zc"
    h.call(f"qlb200_{f.pfx}{'unmrq' if cplx else 'ormrq'}{f.sfx}", side.encode(), ((b"C" if cplx else b"T") if Q.adjoint else b"N"),
           c.m, c.n, k, aptr, f.ld, t.ptr, c.ptr, c.ld)
    return C


class UL:
    """UL factorization object (reference src/ul.jl:51-60): `factors` (U unit upper above the diagonal, L on and
    below it), `ipiv` (1-based int64) and `info`; A[p,:] = U*L."""

    def __init__(self, factors, ipiv, info, handle=None):
        self.factors, self.ipiv, self.info, self.handle = factors, ipiv, int(info), handle

    @property
    def L(self):
        return np.tril(self.factors) if not _is_torch(self.factors) else self.factors.tril()

    @property
    def U(self):
        if _is_torch(self.factors):
            import torch

            return self.factors.triu(1) + torch.eye(*self.factors.shape, dtype=self.factors.dtype, device=self.factors.device)
        return np.triu(self.factors, 1) + np.eye(*self.factors.shape, dtype=self.factors.dtype)

    @property
    def p(self):
        """Row permutation (0-based) such that A[p,:] = U*L (src/ul.jl:317-349 ipiv2perm, reverse order)."""
        m = self.factors.shape[0]
        perm = np.arange(m)
        ip = np.asarray(self.ipiv.cpu() if _is_torch(self.ipiv) else self.ipiv)
        for i in range(len(ip) - 1, -1, -1):
            j = int(ip[i]) - 1
            perm[i], perm[j] = perm[j], perm[i]
        return perm

    def solve(self, b):
        r"""F \ b for square A (src/ul.jl:390-393)."""
        X = _copy_colmajor(b)
        f, bm = _Mat(self.factors, "A(const)"), _Mat(X, "B")
        if f.m != f.n:
            raise DimensionMismatch("UL solve needs a square factorization")
        if bm.m != f.m:
            raise DimensionMismatch(f"arguments must have the same number of rows; has {f.m} and {bm.m}")
        h = _handle(self.handle, f)
        ip = self.ipiv
        iptr = ip.data_ptr() if _is_torch(ip) else ip.ctypes.data
        h.call(f"qlb200_dulsolve{f.sfx}", f.n, bm.n, f.ptr, f.ld, iptr, bm.ptr, bm.ld)
        return X


def ul_(A, pivot=True, check=True, handle=None) -> UL:
    """ul!(A, Val(pivot); check) (src/ul.jl:146-196), Float64."""
    a = _Mat(A)
    if a.pfx != "d":
        raise TypeError("ul_: Float64 only")
    h = _handle(handle, a)
    k = min(a.m, a.n)
    if a.dev:
        import torch

        ipiv = torch.zeros(k, dtype=torch.int64, device=A.device)
        dinfo = torch.zeros(1, dtype=torch.int32, device=A.device)
        h.call("qlb200_dulfact_dev", a.m, a.n, a.ptr, a.ld, ipiv.data_ptr(), int(bool(pivot)), dinfo.data_ptr())
        info = int(dinfo.item())
    else:
        ipiv = np.zeros(k, dtype=np.int64)
        cinfo = _c.c_int32(0)
        h.call("qlb200_dulfact", a.m, a.n, a.ptr, a.ld, ipiv.ctypes.data, int(bool(pivot)), _c.addressof(cinfo))
        info = int(cinfo.value)
    if check and info != 0:          # _checknonsingular (src/ul.jl:193)
        raise SingularException(info)
    return UL(A, ipiv, info, handle)


def ul(A, pivot=True, check=True, handle=None) -> UL:
    """ul(A): copy (promoting integers like ultype, src/ul.jl:198-210), then ul!."""
    if not _is_torch(A):
        A = np.asarray(A)
        if A.dtype != np.float64:
            A = A.astype(np.float64)
    return ul_(_copy_colmajor(A), pivot, check, handle)


def _ormql(side, Q: QLPackedQ, C):
    f, c = _Mat(Q.factors, "A(const)"), _Mat(C, "C")
    t = _Mat(Q.tau, "tau")
    _same_kind(f, c, t)
    h = _handle(Q.handle, f)
    mA, nA = f.m, f.n
    k = min(mA, nA)
    # the reflectors live in the LAST k columns of the factors (src/ql.jl:331,361: k from nA-mA+1)
    aptr = f.ptr + (nA - k) * f.ld * f.itemsize
    nq = c.m if side == "L" else c.n
    if nq != mA:     # src/ql.jl:326-328,356-358,385-387,414-416
        raise DimensionMismatch(f"matrix A has dimensions ({mA},{nA}) but {'B' if side == 'L' else 'A'} has dimensions ({c.m}, {c.n})")
    cplx = f.pfx in "zc"          # complex: ?unmql; Q.T is the ADJOINT Q' (conj at src/ql.jl:336,366,368,400,426,429)
    name = f"qlb200_{f.pfx}{'unmql' if cplx else 'ormql'}{f.sfx}"
    h.call(name, side.encode(), ((b"C" if cplx else b"T") if Q.adjoint else b"N"), c.m, c.n, k, aptr, f.ld, t.ptr, c.ptr, c.ld)
    return C


def lmul_(Q, B):
    """lmul!(Q,B) / lmul!(Q',B): src/ql.jl:321-347 / 350-377 (QLPackedQ), src/rq.jl:130-137,183-202 (RQPackedQ)."""
    if isinstance(Q, QRPackedQ):
        return _ormqr("L", Q, B)
    return _ormrq("L", Q, B) if isinstance(Q, RQPackedQ) else _ormql("L", Q, B)


def rmul_(C, Q):
    """rmul!(C,Q) / rmul!(C,Q'): src/ql.jl:381-406 / 409-435 (QLPackedQ), src/rq.jl:234-291 (RQPackedQ)."""
    if isinstance(Q, QRPackedQ):
        return _ormqr("R", Q, C)
    return _ormrq("R", Q, C) if isinstance(Q, RQPackedQ) else _ormql("R", Q, C)


def ldiv_(F: QL, B):
    """ldiv!(F::QL, B), square (src/ql.jl:438-447,467)."""
    f, b, t = _Mat(F.factors, "A(const)"), _Mat(B, "B"), _Mat(F.tau, "tau")
    _same_kind(f, b, t)
    h = _handle(F.handle, f)
    if f.m < f.n:    # SURVEY.md 8(f) N2, wide: minimum-norm solution; B is the n-row buffer of src/ql.jl:496-501 (b in its first m rows)
        if f.pfx != "d":
            raise TypeError("wide ldiv_: Float64 only")
        if b.m != f.n:
            raise DimensionMismatch(f"wide ldiv_: B must have {f.n} rows (b in the first {f.m}); has {b.m}")
        h.call(f"qlb200_dqlminnorm{f.sfx}", f.m, f.n, b.n, f.ptr, f.ld, t.ptr, b.ptr, b.ld)
        return B
    if b.m != f.m:   # src/ql.jl:494
        raise DimensionMismatch(f"arguments must have the same number of rows; has {f.m} and {b.m}")
    if f.m > f.n:    # least squares; X = B[:n, :] (src/ql.jl:501), B[n:, :] = residual coordinates
        if f.pfx != "d":
            raise TypeError("tall ldiv_: Float64 only")
        h.call(f"qlb200_dqllstsq{f.sfx}", f.m, f.n, b.n, f.ptr, f.ld, t.ptr, b.ptr, b.ld)
        return B
    h.call(f"qlb200_{f.pfx}qlsolve{f.sfx}", f.n, b.n, f.ptr, f.ld, t.ptr, b.ptr, b.ld)
    return B


# ---- batched (new surface; per-matrix semantics = ql!, lmul!, ldiv!) ---------------------------
def banded_ql_(D, m, n, l, u, handle=None):
    """banded_ql!(L::BandedMatrix) (reference ext/MatrixFactorizationsBandedMatricesExt.jl:18-36) on the band data
    D = bandeddata(L) ((l+u+1) x n, column-major; the lower bandwidth already widened as ql(::BandedMatrix) does, ext:14).
    In place; returns (D, tau)."""
    d = _Mat(D, "D")
    if d.pfx != "d":
        raise TypeError("banded_ql_: Float64 only")
    if d.m != l + u + 1 or d.n != n:
        raise DimensionMismatch(f"band data must be {l + u + 1} x {n}, got {d.m} x {d.n}")
    h = _handle(handle, d)
    tau = _new_like(D, min(m, n))
    h.call(f"qlb200_dgbqlf{d.sfx}", m, n, l, u, d.ptr, d.ld, _Mat(tau, "tau").ptr)
    return D, tau


class _Bat:
    def __init__(self, x, what):
        if _is_torch(x):
            if not x.is_cuda or not x.is_contiguous():
                raise ValueError(f"{what}: need a contiguous CUDA tensor")
            self.ptr, self.dev, self.device_index = x.data_ptr(), True, x.device.index or 0
        else:
            if not isinstance(x, np.ndarray) or not x.flags.c_contiguous:
                raise ValueError(f"{what}: need a C-contiguous numpy array (batch, cols, rows)")
            self.ptr, self.dev, self.device_index = x.ctypes.data, False, 0
        self.shape = tuple(x.shape)
        self.pfx = _pfx(x)

    @property
    def sfx(self):
        return "_dev" if self.dev else ""


def ql_batched_(A3, tau=None, handle=None):
    """`for i: ql!(view(A,:,:,i))` in one call.  A3: (batch, n, n).  Returns (A3, tau[batch, n])."""
    a = _Bat(A3, "A3")
    if len(a.shape) != 3 or a.shape[1] != a.shape[2]:
        raise DimensionMismatch("ql_batched_: expected (batch, n, n)")
    batch, n, _ = a.shape
    if tau is None:
        tau = _new_like(A3, batch * n).reshape(batch, n)
    t = _Bat(tau, "tau")
    h = _handle(handle, a)
    h.call(f"qlb200_{a.pfx}geqlf_batched{a.sfx}", n, a.ptr, n, n * n, t.ptr, n, batch)
    return A3, tau


def _apply_batched(name, trans, A3, tau, B3, handle, info=None):
    a, t, b = _Bat(A3, "A3"), _Bat(tau, "tau"), _Bat(B3, "B3")
    if len({(x.dev, x.pfx) for x in (a, t, b)}) != 1:
        raise TypeError("operands must share element type and residency")
    batch, n, _ = a.shape
    if len(b.shape) != 3 or b.shape[0] != batch or b.shape[2] != n:
        raise DimensionMismatch(f"batched factors {a.shape} vs right-hand sides {b.shape}")
    nrhs = b.shape[1]
    h = _handle(handle, a)
    if trans is None:
        iptr = 0 if info is None else (info.data_ptr() if _is_torch(info) else info.ctypes.data)
        h.call(f"qlb200_{a.pfx}qlsolve_batched{a.sfx}", n, nrhs, a.ptr, n, n * n, t.ptr, n, b.ptr, n, n * nrhs, batch, iptr)
    else:
        h.call(f"qlb200_{a.pfx}ormql_batched{a.sfx}", trans, n, nrhs, a.ptr, n, n * n, t.ptr, n, b.ptr, n, n * nrhs, batch)
    return B3


def lmul_batched_(A3, tau, B3, adjoint=False, handle=None):
    """B_i <- Q_i B_i or Q_i' B_i (src/ql.jl:321-347 / 350-377 per matrix).  B3: (batch, nrhs, n)."""
    return _apply_batched("ormql", b"T" if adjoint else b"N", A3, tau, B3, handle)


def ldiv_batched_(A3, tau, B3, info=None, handle=None):
    """B_i <- L_i^{-1} Q_i' B_i (src/ql.jl:440-447,467 per matrix)."""
    return _apply_batched("qlsolve", None, A3, tau, B3, handle, info)


def ql_apply_batched_(A3, B3, op="T", tau=None, info=None, handle=None):
    """Fused `for i: F = ql!(view(A,:,:,i)); lmul!(F.Q', B_i)` (op "T"; "N": Q B_i; "S": ldiv!(F, B_i)) -- BASELINE configs[2]
    and [4] as one call (qlb200_?geqlf_apply_batched): same kernels and bits as ql_batched_ followed by lmul_batched_ /
    ldiv_batched_, the factors stay in L2 between the two (device) or cross PCIe once (host).  Returns (A3, tau, B3)."""
    a, b = _Bat(A3, "A3"), _Bat(B3, "B3")
    if len(a.shape) != 3 or a.shape[1] != a.shape[2]:
        raise DimensionMismatch("ql_apply_batched_: expected (batch, n, n)")
    batch, n, _ = a.shape
    if len(b.shape) != 3 or b.shape[0] != batch or b.shape[2] != n:
        raise DimensionMismatch(f"batched matrices {a.shape} vs right-hand sides {b.shape}")
    if tau is None:
        tau = _new_like(A3, batch * n).reshape(batch, n)
    t = _Bat(tau, "tau")
    if len({(x.dev, x.pfx) for x in (a, t, b)}) != 1:
        raise TypeError("operands must share element type and residency")
    nrhs = b.shape[1]
    h = _handle(handle, a)
    iptr = 0 if info is None else (info.data_ptr() if _is_torch(info) else info.ctypes.data)
    h.call(f"qlb200_{a.pfx}geqlf_apply_batched{a.sfx}", op.encode(), n, nrhs, a.ptr, n, n * n, t.ptr, n, b.ptr, n, n * nrhs, batch, iptr)
    return A3, tau, B3


def ql_batched_mg_(A3, tau=None, handles=None):
    """Host arrays, several GPUs of one box: matrices are sharded by contiguous index ranges over `handles`
    (one per device), no cross-GPU communication; qlb200_?geqlf_batched_mg."""
    a = _Bat(A3, "A3")
    if a.dev:
        raise TypeError("ql_batched_mg_ takes host (NumPy) arrays")
    if len(a.shape) != 3 or a.shape[1] != a.shape[2]:
        raise DimensionMismatch("ql_batched_mg_: expected (batch, n, n)")
    batch, n, _ = a.shape
    if tau is None:
        tau = np.empty((batch, n), dtype=A3.dtype)
    t = _Bat(tau, "tau")
    if handles is None:              # one handle per visible device
        import torch

        handles = [default_handle(d) for d in range(max(1, torch.cuda.device_count()))]
    _lib.call_mg(list(handles), f"qlb200_{a.pfx}geqlf_batched_mg", n, a.ptr, n, n * n, t.ptr, n, batch)
    return A3, tau
