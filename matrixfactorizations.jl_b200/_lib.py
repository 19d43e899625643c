"""ctypes binding of libqlb200.so -- the C-ABI declared in include/qlb200.h.

This is the same binding a Julia `ccall` makes: plain pointers and int64 sizes.
There is no CPU fallback: if the shared library is missing or no sm_100 device is present every
call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char, c_char_p, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqlb200.so")
_LIB = None


class QLB200Error(RuntimeError):
    """CUDA/runtime failure reported by libqlb200 (info >= 1000)."""


class ArgumentError(ValueError):
    """info == -i for a non-dimension argument (Julia: ArgumentError)."""


class DimensionMismatch(ValueError):
    """Dimension check failed (reference: DimensionMismatch, src/ql.jl:326-328,356-358)."""


class SingularException(ArithmeticError):
    """Zero diagonal entry met in a triangular solve / UL (reference: SingularException)."""

    def __init__(self, info):
        super().__init__(f"SingularException({info})")
        self.info = info


_I, _P = c_int64, c_void_p

# name -> argtypes (after the handle).  Mirrors include/qlb200.h one to one.
_FACT = [_I, _I, _P, _I, _P]
_ORM = [c_char, c_char, _I, _I, _I, _P, _I, _P, _P, _I]
_SOLVE = [_I, _I, _P, _I, _P, _P, _I]
_FACT_B = [_I, _P, _I, _I, _P, _I, _I]
_ORM_B = [c_char, _I, _I, _P, _I, _I, _P, _I, _P, _I, _I, _I]
_SOLVE_B = [_I, _I, _P, _I, _I, _P, _I, _P, _I, _I, _I, _P]
_UL = [_I, _I, _P, _I, _P, c_int, _P]
_ULS = [_I, _I, _P, _I, _P, _P, _I]
SIGNATURES = {}
for _sfx in ("", "_dev"):
    SIGNATURES[f"qlb200_dgeqlf{_sfx}"] = _FACT
    SIGNATURES[f"qlb200_dgerqf{_sfx}"] = _FACT
    SIGNATURES[f"qlb200_dormql{_sfx}"] = _ORM
    SIGNATURES[f"qlb200_dqlsolve{_sfx}"] = _SOLVE
    SIGNATURES[f"qlb200_dormrq{_sfx}"] = _ORM
    SIGNATURES[f"qlb200_dgeqrf{_sfx}"] = _FACT
    SIGNATURES[f"qlb200_dormqr{_sfx}"] = _ORM
    SIGNATURES[f"qlb200_dqllstsq{_sfx}"] = [_I, _I, _I, _P, _I, _P, _P, _I]
    SIGNATURES[f"qlb200_dqrsolve{_sfx}"] = [_I, _I, _I, _P, _I, _P, _P, _I]
    SIGNATURES[f"qlb200_dqlminnorm{_sfx}"] = [_I, _I, _I, _P, _I, _P, _P, _I]
    SIGNATURES[f"qlb200_dgbqlf{_sfx}"] = [_I, _I, _I, _I, _P, _I, _P]
    SIGNATURES[f"qlb200_dulfact{_sfx}"] = _UL
    SIGNATURES[f"qlb200_dulsolve{_sfx}"] = _ULS
    # Float32 (promoted), ComplexF64 (complex.cu) and ComplexF32 (promoted) single-matrix QL path
    SIGNATURES[f"qlb200_sgeqlf{_sfx}"] = _FACT
    SIGNATURES[f"qlb200_sormql{_sfx}"] = _ORM
    SIGNATURES[f"qlb200_sqlsolve{_sfx}"] = _SOLVE
    SIGNATURES[f"qlb200_sgerqf{_sfx}"] = _FACT
    SIGNATURES[f"qlb200_sormrq{_sfx}"] = _ORM
    for _t in "zc":
        SIGNATURES[f"qlb200_{_t}geqlf{_sfx}"] = _FACT
        SIGNATURES[f"qlb200_{_t}unmql{_sfx}"] = _ORM
        SIGNATURES[f"qlb200_{_t}qlsolve{_sfx}"] = _SOLVE
        SIGNATURES[f"qlb200_{_t}gerqf{_sfx}"] = _FACT
        SIGNATURES[f"qlb200_{_t}unmrq{_sfx}"] = _ORM
    for _t in "ds":
        SIGNATURES[f"qlb200_{_t}geqlf_batched{_sfx}"] = _FACT_B
        SIGNATURES[f"qlb200_{_t}ormql_batched{_sfx}"] = _ORM_B
        SIGNATURES[f"qlb200_{_t}qlsolve_batched{_sfx}"] = _SOLVE_B
        SIGNATURES[f"qlb200_{_t}geqlf_apply_batched{_sfx}"] = [c_char, _I, _I, _P, _I, _I, _P, _I, _P, _I, _I, _I, _P]
# multi-GPU forms: the first argument is an array of handles + its length instead of one handle
MG_SIGNATURES = {}
for _t in "ds":
    MG_SIGNATURES[f"qlb200_{_t}geqlf_batched_mg"] = _FACT_B
    MG_SIGNATURES[f"qlb200_{_t}ormql_batched_mg"] = _ORM_B
    MG_SIGNATURES[f"qlb200_{_t}qlsolve_batched_mg"] = _SOLVE_B
SIGNATURES["qlb200_host_register"] = [_P, _I]
SIGNATURES["qlb200_host_unregister"] = [_P]
SIGNATURES["qlb200_dgemm_dev"] = [c_char, c_char, _I, _I, _I, ctypes.c_double, _P, _I, _P, _I, ctypes.c_double, _P, _I]
HANDLE_FUNCS = ["qlb200_version", "qlb200_create", "qlb200_destroy", "qlb200_set_stream", "qlb200_sync",
                "qlb200_last_error", "qlb200_launch_count", "qlb200_set_option", "qlb200_reset_stream",
                "qlb200_profile_collect", "qlb200_upload_rate"]


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise QLB200Error(f"{LIB_PATH} is missing: build it with `make -C {_HERE}/csrc` "
                              "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.qlb200_version.restype = c_int
        L.qlb200_create.argtypes = [POINTER(c_void_p), c_int]
        L.qlb200_destroy.argtypes = [c_void_p]
        L.qlb200_set_stream.argtypes = [c_void_p, c_void_p]
        L.qlb200_reset_stream.argtypes = [c_void_p]
        L.qlb200_profile_collect.argtypes = [c_void_p, POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int64)]
        L.qlb200_sync.argtypes = [c_void_p]
        L.qlb200_last_error.argtypes = [c_void_p]
        L.qlb200_last_error.restype = c_char_p
        L.qlb200_launch_count.argtypes = [c_void_p]
        L.qlb200_launch_count.restype = c_int64
        L.qlb200_set_option.argtypes = [c_void_p, c_char_p, c_int64]
        L.qlb200_upload_rate.argtypes = [c_void_p]
        L.qlb200_upload_rate.restype = ctypes.c_double
        _LIB = L
    return _LIB


_BOUND = {}


def fn(name: str):
    """The C entry point `name` with its argtypes set (bound on first use)."""
    f = _BOUND.get(name)
    if f is None:
        if name not in SIGNATURES:        # e.g. a complex / Float32 variant this library does not provide
            raise TypeError(f"{name}: no such entry point in libqlb200 (element type not supported for this operation; "
                            "see include/qlb200.h)")
        f = getattr(lib(), name)          # AttributeError if the library does not export it
        f.argtypes = [c_void_p] + SIGNATURES[name]
        f.restype = c_int
        _BOUND[name] = f
    return f


def fn_mg(name: str):
    f = _BOUND.get(name)
    if f is None:
        f = getattr(lib(), name)
        f.argtypes = [POINTER(c_void_p), c_int] + MG_SIGNATURES[name]
        f.restype = c_int
        _BOUND[name] = f
    return f


def call_mg(handles, name: str, *args):
    """Call a `_mg` entry point with a list of Handle objects (one per device)."""
    arr = (c_void_p * len(handles))(*[h._h for h in handles])
    info = fn_mg(name)(arr, len(handles), *args)
    handles[0].check(info)
    return info


class Handle:
    """qlb200_handle: one CUDA device, its stream and workspaces."""

    def __init__(self, device: int = 0):
        self._h = c_void_p()
        rc = lib().qlb200_create(ctypes.byref(self._h), int(device))
        if rc != 0:
            self._h = c_void_p()
            raise QLB200Error(f"qlb200_create(device={device}) failed: info={rc}"
                              + (" (CUDA error %d)" % (rc - 1000) if rc >= 1000 else ""))
        self.device = int(device)

    def close(self):
        if self._h:
            lib().qlb200_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(lib().qlb200_sync(self._h))

    def set_stream(self, cuda_stream_ptr):
        self.check(lib().qlb200_set_stream(self._h, c_void_p(cuda_stream_ptr or 0)))

    def reset_stream(self):
        self.check(lib().qlb200_reset_stream(self._h))

    def set_option(self, name: str, value: int):
        self.check(lib().qlb200_set_option(self._h, name.encode(), int(value)))

    def profile_collect(self):
        """(device ms, algorithmic flops, launches) of the bracketed GEMM launches since the last call."""
        ms, fl, n = ctypes.c_double(0), ctypes.c_double(0), c_int64(0)
        self.check(lib().qlb200_profile_collect(self._h, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n)))
        return ms.value, fl.value, n.value

    @property
    def launches(self) -> int:
        return int(lib().qlb200_launch_count(self._h))

    def upload_rate(self) -> float:
        """GB/s the first upload piece of the last large host-pointer ql! saw (0.0: none yet)."""
        return float(lib().qlb200_upload_rate(self._h))

    def pin(self, a) -> None:
        """Page-lock the NumPy array `a` (cudaHostRegister): host-pointer calls on it copy at PCIe speed and a repeated
        large ql_ replays as one CUDA graph.  Undo with unpin(a) before the array is freed."""
        self.call("qlb200_host_register", a.ctypes.data, a.nbytes)

    def unpin(self, a) -> None:
        self.call("qlb200_host_unregister", a.ctypes.data)

    def last_error(self) -> str:
        return lib().qlb200_last_error(self._h).decode(errors="replace")

    def check(self, info: int, dim_args=()):
        """Map the LAPACK-style info to the exception types the reference throws."""
        if info == 0:
            return
        msg = self.last_error()
        if info >= 1000:
            raise QLB200Error(f"CUDA error {info - 1000}: {msg}")
        if info > 0:
            raise SingularException(info)
        if (-info) in dim_args or "DimensionMismatch" in msg:
            raise DimensionMismatch(f"argument {-info}: {msg}")
        raise ArgumentError(f"argument {-info}: {msg}")

    def call(self, name: str, *args, dim_args=()):
        info = fn(name)(self._h, *args)
        self.check(info, dim_args)
        return info

    def call_raw(self, name: str, *args) -> int:
        return fn(name)(self._h, *args)


_DEFAULT = {}


def default_handle(device: int = 0) -> Handle:
    h = _DEFAULT.get(device)
    if h is None:
        h = _DEFAULT[device] = Handle(device)
    return h
