"""Small fixed workloads for ncu (run under gpurun).  usage: prof_case.py <case> [n]
 cases: geqlf (one blocked ql! of n x n), gemms (one C'V and one rank-128 update at n), batched (32x32 f64)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
q = load_package()
h = q.Handle(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    h.set_option(k, int(v))
cm = lambda r, c: torch.randn((c, r), dtype=torch.float64, device="cuda").T
if case == "geqlf":
    A = cm(n, n)
    tau = torch.empty(n, dtype=torch.float64, device="cuda")
    h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())
elif case == "gemms":
    C, V, W = cm(n, n), cm(n, 128), cm(n, 128)
    h.call("qlb200_dgemm_dev", b"T", b"N", n, 128, n, 1.0, C.data_ptr(), C.stride(1), V.data_ptr(), V.stride(1), 0.0,
           W.data_ptr(), W.stride(1))
    h.call("qlb200_dgemm_dev", b"N", b"T", n, n, 128, -1.0, V.data_ptr(), V.stride(1), W.data_ptr(), W.stride(1), 1.0,
           C.data_ptr(), C.stride(1))
elif case == "batched":
    A = torch.randn(n, 32, 32, dtype=torch.float64, device="cuda")
    tau = torch.empty(n, 32, dtype=torch.float64, device="cuda")
    B = torch.randn(n, 8, 32, dtype=torch.float64, device="cuda")
    q.ql_batched_(A, tau, handle=h)
    q.lmul_batched_(A, tau, B, adjoint=True, handle=h)
torch.cuda.synchronize()
print("done", case, n, "launches", h.launches)
if case == "ormql":
    A = cm(n, n)
    tau = torch.empty(n, dtype=torch.float64, device="cuda")
    h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())
    B = cm(n, 8)
    torch.cuda.synchronize()
    l0 = h.launches
    h.call("qlb200_dormql_dev", b"L", b"T", n, 8, n, A.data_ptr(), A.stride(1), tau.data_ptr(), B.data_ptr(), B.stride(1))
    torch.cuda.synchronize()
    print("ormql launches", h.launches - l0)
if case == "ul":
    A = cm(n, n)
    ipiv = torch.zeros(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    h.call("qlb200_dulfact_dev", n, n, A.data_ptr(), A.stride(1), ipiv.data_ptr(), 1, info.data_ptr())
    torch.cuda.synchronize()
    print("ul done", h.launches)
if case == "batched16":
    A = torch.randn(n, 16, 16, dtype=torch.float32, device="cuda")
    tau = torch.empty(n, 16, dtype=torch.float32, device="cuda")
    B = torch.randn(n, 8, 16, dtype=torch.float32, device="cuda")
    q.ql_batched_(A, tau, handle=h)
    q.ldiv_batched_(A, tau, B, handle=h)
    torch.cuda.synchronize()
    print("batched16 done", h.launches)
if case == "strips":
    # strips of different widths / heights: per-column cost vs fixed cost of the strip kernel
    for (m, w) in [(16384, 16), (16384, 8), (16384, 1), (1024, 16), (1024, 1), (16, 16)]:
        A = cm(m, w)
        tau = torch.empty(w, dtype=torch.float64, device="cuda")
        h.set_option("graph", 0)
        h.call("qlb200_dgeqlf_dev", m, w, A.data_ptr(), A.stride(1), tau.data_ptr())
    torch.cuda.synchronize()
    print("strips done")
if case == "bql":
    # batched Float64 QL of order n (PROF_BATCH matrices): the shared-memory WY kernel for n > 32
    nb = int(os.environ.get("PROF_BATCH", "32768"))
    A = torch.randn(nb, n, n, dtype=torch.float64, device="cuda")
    tau = torch.empty(nb, n, dtype=torch.float64, device="cuda")
    q.ql_batched_(A, tau, handle=h)
    torch.cuda.synchronize()
    print("bql done", h.launches)
if case == "zgeqlf":
    A = torch.empty_strided((n, n), (1, n), dtype=torch.complex128, device="cuda").copy_(torch.randn(n, n, dtype=torch.complex128, device="cuda"))
    import time
    for it in range(2):
        B = A.clone(memory_format=torch.preserve_format)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        q.ql_(B, handle=h); torch.cuda.synchronize()
        print("zgeqlf", n, "ms", (time.perf_counter() - t0) * 1e3, "launches", h.launches)
