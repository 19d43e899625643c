"""W = C'V GEMM timing for every tile configuration at several M (run under gpurun)."""
import sys, os, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
cm = lambda r, c: torch.randn((c, r), dtype=torch.float64, device="cuda").T
p, nb = 16384, 128
C = cm(p, 16384); V = cm(p, nb); W = cm(16384, nb)
def t_of(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)
for M in (16256, 12288, 8192, 4096):
    for cfg in (1, 4, 5, 6, 0):
        h.set_option("gemm_cfg", cfg)
        t = t_of(lambda: h.call("qlb200_dgemm_dev", b"T", b"N", M, nb, p, 1.0, C.data_ptr(), C.stride(1), V.data_ptr(), V.stride(1), 0.0, W.data_ptr(), W.stride(1)))
        print(json.dumps({"M": M, "cfg": cfg, "ms": t * 1e3, "tflops": 2.0 * p * M * nb / t / 1e12}), flush=True)
