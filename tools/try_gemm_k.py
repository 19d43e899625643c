"""Rate of the two trailing-update GEMMs (W = C'V and C += V W2') against the block-reflector width K (run under gpurun)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
C = torch.randn((n, n), dtype=torch.float64, device="cuda").T          # column-major n x n
for K in (128, 192, 256, 384, 512):
    V = torch.randn((K, n), dtype=torch.float64, device="cuda").T      # n x K, ld n
    W = torch.zeros((K, n), dtype=torch.float64, device="cuda").T      # n x K
    res = {}
    for name, call in (("W=C'V", lambda: h.call("qlb200_dgemm_dev", b"T", b"N", n, K, n, 1.0, C.data_ptr(), n, V.data_ptr(), n, 0.0, W.data_ptr(), n)),
                       ("C+=VW'", lambda: h.call("qlb200_dgemm_dev", b"N", b"T", n, n, K, 1e-9, V.data_ptr(), n, W.data_ptr(), n, 1.0, C.data_ptr(), n))):
        ts = []
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            if i: ts.append(e0.elapsed_time(e1))
        res[name] = round(2.0 * n * n * K / min(ts) / 1e9, 2)
    print(json.dumps({"n": n, "K": K, "TFLOP/s": res}), flush=True)
