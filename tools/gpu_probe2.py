"""GPU probe for the blocked path: DMMA GEMM on the trailing-update shapes (all tile configs) and
qlb200_dgeqlf_dev at several n / nb.  Writes gpurun_out/probe2.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def ev_time(fn, iters=3, warm=1):
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
    return min(ts)


def main():
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [4096, 8192, 16384]
    out = {}
    q = load_package()
    h = q.Handle(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    h.set_stream(st.cuda_stream)
    cm = lambda r, c: torch.randn((c, r), dtype=torch.float64, device="cuda").T
    p = qd = 16384
    nb = 128
    C = cm(p, qd)
    V = cm(p, nb)
    W = cm(qd, nb)
    for cfg in (-1, 0, 1, 3):
        h.set_option("gemm_cfg", cfg)
        # W = C' V : M = q, N = nb, K = p  (TN)
        t = ev_time(lambda: h.call("qlb200_dgemm_dev", b"T", b"N", qd, nb, p, 1.0, C.data_ptr(), C.stride(1), V.data_ptr(),
                                   V.stride(1), 0.0, W.data_ptr(), W.stride(1)))
        out[f"gemm_CtV_cfg{cfg}_tflops"] = 2.0 * p * qd * nb / t / 1e12
        # C -= V W' : M = p, N = q, K = nb (NT)
        t = ev_time(lambda: h.call("qlb200_dgemm_dev", b"N", b"T", p, qd, nb, -1.0, V.data_ptr(), V.stride(1), W.data_ptr(),
                                   W.stride(1), 1.0, C.data_ptr(), C.stride(1)))
        out[f"gemm_rankk_cfg{cfg}_tflops"] = 2.0 * p * qd * nb / t / 1e12
        # square 8192 NN
        A2, B2, C2 = cm(8192, 8192), cm(8192, 8192), cm(8192, 8192)
        t = ev_time(lambda: h.call("qlb200_dgemm_dev", b"N", b"N", 8192, 8192, 8192, 1.0, A2.data_ptr(), A2.stride(1),
                                   B2.data_ptr(), B2.stride(1), 0.0, C2.data_ptr(), C2.stride(1)))
        out[f"gemm_sq8192_cfg{cfg}_tflops"] = 2.0 * 8192 ** 3 / t / 1e12
        del A2, B2, C2
    h.set_option("gemm_cfg", -1)
    del C, V, W
    for n in sizes:
        A0 = cm(n, n)
        A = torch.empty_like(A0.T).T
        tau = torch.empty(n, dtype=torch.float64, device="cuda")
        for (nbv, la, rpt) in ((128, 0, 4), (128, 0, 1), (96, 0, 4)):
            h.set_option("nb", nbv)
            h.set_option("lookahead", la)
            h.set_option("strip_rpt", rpt)

            def run():
                A.copy_(A0)
                h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())

            tcopy = ev_time(lambda: A.copy_(A0))
            l0 = h.launches
            t = ev_time(run, iters=2, warm=1) - tcopy
            out[f"geqlf_n{n}_nb{nbv}_la{la}_rpt{rpt}"] = {"s": t, "tflops": 4.0 / 3.0 * n ** 3 / t / 1e12,
                                                        "launches": (h.launches - l0) // 3}
        del A0, A, tau
    h.set_option("nb", 128)
    h.set_option("lookahead", 0)
    h.set_option("strip_rpt", 4)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe2.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
