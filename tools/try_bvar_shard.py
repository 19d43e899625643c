"""n = 32 Float64 batched QL at shard sizes: single kernel (bvar 2) vs two-phase split (bvar 3) (run under gpurun)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
n = 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for batch in (16384, 32768, 49152, 65536, 131072):
    A0 = torch.randn((batch, n, n), dtype=torch.float64, device="cuda")
    A = torch.empty_like(A0); tau = torch.empty((batch, n), dtype=torch.float64, device="cuda")
    out = {}
    for bvar in (0, 2, 3):
        h.set_option("bvar", bvar)
        ts = []
        for i in range(8):
            A.copy_(A0); flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.call("qlb200_dgeqlf_batched_dev", n, A.data_ptr(), n, n * n, tau.data_ptr(), n, batch)
            e1.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(e0.elapsed_time(e1))
        out[bvar] = round(sorted(ts)[len(ts) // 2], 4)
    print(json.dumps({"batch": batch, "ms by bvar": out}), flush=True)
h.set_option("bvar", 0)
