"""Timings of the non-headline single-matrix rows (run under gpurun): UL, ormql (lmul!), qlsolve, rq."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
cm = lambda r, c: torch.randn((c, r), dtype=torch.float64, device="cuda").T
def t_of(fn, reset=None, iters=3):
    ts = []
    for i in range(iters + 1):
        if reset: reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i: ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)
for n in (1024, 4096, 8192, 16384):
    A0 = cm(n, n); A = cm(n, n)
    ipiv = torch.zeros(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    t = t_of(lambda: h.call("qlb200_dulfact_dev", n, n, A.data_ptr(), A.stride(1), ipiv.data_ptr(), 1, info.data_ptr()), lambda: A.T.copy_(A0.T))
    print(json.dumps({"op": "ulfact", "n": n, "ms": t * 1e3, "tflops": 2 / 3 * n ** 3 / t / 1e12}), flush=True)
    tau = torch.empty(n, dtype=torch.float64, device="cuda")
    A.T.copy_(A0.T)
    h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())
    for nb in (1, 8, 128, n):
        if nb == n and n > 8192: continue
        B = cm(n, nb)
        t = t_of(lambda: h.call("qlb200_dormql_dev", b"L", b"T", n, nb, n, A.data_ptr(), A.stride(1), tau.data_ptr(), B.data_ptr(), B.stride(1)))
        print(json.dumps({"op": "ormql L,T", "n": n, "nB": nb, "ms": t * 1e3, "tflops": 2.0 * n * n * nb / t / 1e12, "V_stream_gbs": n * n * 8 / t / 1e9}), flush=True)
        t = t_of(lambda: h.call("qlb200_dqlsolve_dev", n, nb, A.data_ptr(), A.stride(1), tau.data_ptr(), B.data_ptr(), B.stride(1)))
        print(json.dumps({"op": "qlsolve", "n": n, "nB": nb, "ms": t * 1e3, "tflops": 3.0 * n * n * nb / t / 1e12}), flush=True)
        del B
    del A0, A
