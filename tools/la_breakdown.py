"""ql! n=16384: total time vs the summed duration of the bracketed trailing-update GEMMs, per lookahead mode (run under gpurun)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
for kv in sys.argv[2:]:
    k, v = kv.split("="); h.set_option(k, int(v))
A0 = torch.randn((n, n), dtype=torch.float64, device="cuda").T
A = torch.empty((n, n), dtype=torch.float64, device="cuda").T
tau = torch.empty(n, dtype=torch.float64, device="cuda")
for la in (0, 2):
    h.set_option("lookahead", la)
    for prof in (0, 1):
        h.set_option("profile", prof)
        ts = []
        for i in range(3):
            A.copy_(A0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr()); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            if prof:
                ms, fl, cnt = h.profile_collect()
        out = {"n": n, "lookahead": la, "profile": prof, "ms": min(ts[1:])}
        if prof:
            out.update({"gemm_ms": ms, "gemm_tflops": fl / ms / 1e9, "gemm_launches": cnt})
        print(json.dumps(out))
h.set_option("profile", 0)
