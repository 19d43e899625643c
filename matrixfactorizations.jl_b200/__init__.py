"""qlb200: B200-native (sm_100a) Householder QL hot path behind MatrixFactorizations.jl's API.

The directory name carries a dot, so it is not importable by name; load it with
`__graft_entry__.load_package()` (imports it as module `qlb200`).
"""
from ._lib import (ArgumentError, DimensionMismatch, Handle, QLB200Error, SingularException,  # noqa: F401
                   LIB_PATH, SIGNATURES, MG_SIGNATURES, HANDLE_FUNCS, default_handle, lib, call_mg)
from .ql import (QL, QLPackedQ, RQPackedQ, QRPackedQ, UL, ql, ql_, rq_, qr_, qr_solve, ul, ul_, lmul_, rmul_, ldiv_, ql_batched_,  # noqa: F401
                 lmul_batched_, ldiv_batched_, ql_batched_mg_, ql_apply_batched_, banded_ql_)
from . import shard  # noqa: F401
