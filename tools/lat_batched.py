"""Latency structure of the batched n=32 f64 kernel: time per warp-batch with 1, 2, 3 resident CTAs per SM (run under gpurun)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package()
h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
n = 32
out = {}
bvars = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "2").split(",")]
for bvar in bvars:
    h.set_option("bvar", bvar)
    for ctas_per_sm in (1, 2, 3, 4):
        grid = 148 * ctas_per_sm
        batch = grid * 4 * 2 * 24          # 24 warp-batches per warp
        A0 = torch.randn(batch, n, n, dtype=torch.float64, device="cuda"); A = A0.clone()
        tau = torch.empty(batch, n, dtype=torch.float64, device="cuda")
        h.set_option("bgrid", grid)
        ts = []
        for i in range(6):
            A.copy_(A0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); q.ql_batched_(A, tau, handle=h); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[f"bvar{bvar}_ctas{ctas_per_sm}"] = {"ms": ms, "us_per_warp_batch": ms * 1e3 / 24, "matrices_per_s": batch / ms * 1e3}
h.set_option("bgrid", 0)
print(json.dumps(out, indent=1))
