"""Batched-kernel timing + parity spot check (run under gpurun).
usage: bench_batched.py [n=32] [dtype=f64] [batch=262144] [bvar=0,1,...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import cport  # noqa: E402

kw = dict(a.split("=") for a in sys.argv[1:])
n = int(kw.get("n", 32))
dt = torch.float64 if kw.get("dtype", "f64") == "f64" else torch.float32
batch = int(kw.get("batch", 262144))
bvars = [int(x) for x in kw.get("bvar", "0").split(",")]
nrhs = 8
q = load_package()
h = q.Handle(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1234321)
A0 = torch.randn(batch, n, n, dtype=dt, device="cuda", generator=g)
A = A0.clone()
tau = torch.empty(batch, n, dtype=dt, device="cuda")
B0 = torch.randn(batch, nrhs, n, dtype=dt, device="cuda", generator=g)
B = B0.clone()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
es = A.element_size()
sub = 512
Fo, tauo = cport.geql2_batched(A0[:sub].cpu().numpy())
eps = np.finfo(np.float64 if dt == torch.float64 else np.float32).eps


def timed(fn, iters=7):
    ts = []
    for i in range(iters + 2):
        A.copy_(A0)
        B.copy_(B0)
        flush.fill_(i & 255)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts)), min(ts)


out = {}
for bv in bvars:
    h.set_option("bvar", bv)
    med, best = timed(lambda: q.ql_batched_(A, tau, handle=h))
    err = float(np.abs(A[:sub].cpu().numpy() - Fo).max())
    terr = float(np.abs(tau[:sub].cpu().numpy() - tauo).max())
    r = {"geqlf_ms": med * 1e3, "best_ms": best * 1e3, "matrices_per_s": batch / med, "gbs": batch * (2 * n * n + n) * es / med / 1e9,
         "max_err_factors": err, "max_err_tau": terr, "tol": float(50 * n * eps * n)}
    # the factors are in A now: time the apply kernels on them
    F = A.clone()
    med, best = timed(lambda: q.lmul_batched_(F, tau, B, adjoint=True, handle=h))
    r["ormqlT_ms"] = med * 1e3
    r["ormqlT_gbs"] = batch * (n * n + n + 2 * n * nrhs) * es / med / 1e9
    med, best = timed(lambda: q.ldiv_batched_(F, tau, B, handle=h))
    r["qlsolve_ms"] = med * 1e3
    r["qlsolve_gbs"] = batch * (n * n + n + 2 * n * nrhs) * es / med / 1e9
    out[f"n{n}_{kw.get('dtype', 'f64')}_bvar{bv}"] = r
print(json.dumps(out, indent=1))
