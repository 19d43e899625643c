"""One-off GPU probe (run under gpurun): measured FP64 DGEMM peak (cuBLAS via torch, yardstick only),
HBM copy bandwidth, and batched-kernel timings.  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def ev_time(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts), float(np.median(ts))


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    q = load_package()
    h = q.Handle(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    h.set_stream(st.cuda_stream)
    # FP64 DGEMM yardstick
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        c = torch.empty_like(a)
        best, med = ev_time(lambda: torch.matmul(a, b, out=c), iters=5)
        out[f"cublas_dgemm_{n}_tflops_best"] = 2 * n ** 3 / best / 1e12
        out[f"cublas_dgemm_{n}_tflops_median"] = 2 * n ** 3 / med / 1e12
        del a, b, c
    # skinny shapes of the trailing update (K = 128)
    p = qd = 16384
    V = torch.randn(128, p, dtype=torch.float64, device="cuda")       # column-major p x 128
    C = torch.randn(qd, p, dtype=torch.float64, device="cuda")        # column-major p x q
    W = torch.empty(qd, 128, dtype=torch.float64, device="cuda")
    best, med = ev_time(lambda: torch.matmul(C, V.T, out=W), iters=5)   # W = C^T V  (q x 128)
    out["cublas_CtV_16384x128_tflops"] = 2 * p * qd * 128 / best / 1e12
    best, med = ev_time(lambda: torch.addmm(C, W, V, alpha=-1.0, out=C), iters=5)   # C -= W V^T (row-major view)
    out["cublas_rankk_16384x128_tflops"] = 2 * p * qd * 128 / best / 1e12
    del V, C, W
    # HBM copy
    x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    best, med = ev_time(lambda: y.copy_(x), iters=10)
    out["hbm_copy_gbs"] = 2 * x.numel() / best / 1e9
    del x, y
    # batched kernels
    for (n, dt, batch, nrhs) in [(32, torch.float64, 262144, 8), (16, torch.float32, 1048576, 8), (16, torch.float64, 524288, 8),
                                 (8, torch.float64, 1048576, 8), (32, torch.float32, 262144, 8)]:
        A0 = torch.randn(batch, n, n, dtype=dt, device="cuda")
        A = A0.clone()
        tau = torch.empty(batch, n, dtype=dt, device="cuda")
        B = torch.randn(batch, nrhs, n, dtype=dt, device="cuda")
        es = A.element_size()

        def fact():
            A.copy_(A0)
            q.ql_batched_(A, tau, handle=h)

        tcopy, _ = ev_time(lambda: A.copy_(A0), iters=5)
        best, med = ev_time(fact, iters=5)
        tk = best - tcopy
        key = f"batched_n{n}_{'f64' if dt == torch.float64 else 'f32'}"
        out[key] = {"batch": batch, "geqlf_s": tk, "matrices_per_s": batch / tk,
                    "gbs": batch * (2 * n * n + n) * es / tk / 1e9}
        q.ql_batched_(A, tau, handle=h)
        best, med = ev_time(lambda: q.lmul_batched_(A, tau, B, adjoint=True, handle=h), iters=5)
        out[key]["ormql_T_s"] = best
        out[key]["ormql_gbs"] = batch * (n * n + n + 2 * n * nrhs) * es / best / 1e9
        best, med = ev_time(lambda: q.ldiv_batched_(A, tau, B, handle=h), iters=5)
        out[key]["qlsolve_s"] = best
        out[key]["qlsolve_gbs"] = batch * (n * n + n + 2 * n * nrhs) * es / best / 1e9
        del A0, A, tau, B
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
