"""One table of device-side timings for every entry point family (run under gpurun) -> profiles/r01_sweep.md"""
import os, sys, json
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
cm = lambda r, c: torch.randn((c, r), dtype=torch.float64, device="cuda").T
def t_of(fn, reset=None, iters=3):
    ts = []
    for i in range(iters + 2):
        if reset: reset()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)
out = ["# Device-side timings, one B200 (tools/sweep_all.py; CUDA events, best of 3 after 2 warm-ups)", "",
       "## Single matrix, Float64", "", "| n | ql! ms | ql! TFLOP/s | lmul!(Q',B) nB=8 ms | F\\\\b nB=1 ms | ul! ms | ul! TFLOP/s |", "|---|---|---|---|---|---|---|"]
for n in (1024, 2048, 4096, 8192, 12288, 16384):
    A0 = cm(n, n); A = cm(n, n); tau = torch.empty(n, dtype=torch.float64, device="cuda")
    tq = t_of(lambda: h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr()), lambda: A.T.copy_(A0.T))
    B = cm(n, 8); b1 = cm(n, 1)
    tl = t_of(lambda: h.call("qlb200_dormql_dev", b"L", b"T", n, 8, n, A.data_ptr(), A.stride(1), tau.data_ptr(), B.data_ptr(), B.stride(1)))
    ts_ = t_of(lambda: h.call("qlb200_dqlsolve_dev", n, 1, A.data_ptr(), A.stride(1), tau.data_ptr(), b1.data_ptr(), b1.stride(1) if n > 1 else n))
    ipiv = torch.zeros(n, dtype=torch.int64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
    tu = t_of(lambda: h.call("qlb200_dulfact_dev", n, n, A.data_ptr(), A.stride(1), ipiv.data_ptr(), 1, info.data_ptr()), lambda: A.T.copy_(A0.T))
    out.append(f"| {n} | {tq*1e3:.2f} | {4/3*n**3/tq/1e12:.2f} | {tl*1e3:.2f} | {ts_*1e3:.2f} | {tu*1e3:.2f} | {2/3*n**3/tu/1e12:.2f} |")
    del A0, A, B, b1
out += ["", "## Batched (HBM peak 6549.8 GB/s measured)", "", "| n | type | batch | ql ms | ql M matrices/s | ql % HBM | lmul(Q',B) nB=8 % HBM | ldiv nB=8 % HBM |", "|---|---|---|---|---|---|---|---|"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (n, dt, batch) in [(32, torch.float64, 262144), (16, torch.float64, 524288), (8, torch.float64, 1048576), (32, torch.float32, 262144), (16, torch.float32, 1048576), (8, torch.float32, 2097152)]:
    es = 8 if dt == torch.float64 else 4
    A0 = torch.randn(batch, n, n, dtype=dt, device="cuda"); A = A0.clone()
    tau = torch.empty(batch, n, dtype=dt, device="cuda")
    B0 = torch.randn(batch, 8, n, dtype=dt, device="cuda"); B = B0.clone()
    def rs():
        A.copy_(A0); B.copy_(B0); flush.fill_(1)
    tf = t_of(lambda: q.ql_batched_(A, tau, handle=h), rs)
    q.ql_batched_(A, tau, handle=h)
    F = A.clone()
    def rsb():
        B.copy_(B0); flush.fill_(1)
    tl = t_of(lambda: q.lmul_batched_(F, tau, B, adjoint=True, handle=h), rsb)
    ts_ = t_of(lambda: q.ldiv_batched_(F, tau, B, handle=h), rsb)
    by_f, by_a = (2 * n * n + n) * es, (n * n + n + 2 * n * 8) * es
    out.append(f"| {n} | {'f64' if es == 8 else 'f32'} | {batch} | {tf*1e3:.3f} | {batch/tf/1e6:.0f} | {100*batch*by_f/tf/6549.8e9:.0f} | {100*batch*by_a/tl/6549.8e9:.0f} | {100*batch*by_a/ts_/6549.8e9:.0f} |")
    del A0, A, tau, B0, B, F
open(os.path.join(ROOT, "gpurun_out", "r01_sweep.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
