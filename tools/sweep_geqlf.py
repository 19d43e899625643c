"""Option sweep for the blocked ql! (run under gpurun).  usage: sweep_geqlf.py n key=v1,v2 ..."""
import itertools
import json
import sys
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

n = int(sys.argv[1])
grid = {}
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    grid[k] = [int(x) for x in v.split(",")]
q = load_package()
h = q.Handle(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)
A0 = torch.randn((n, n), dtype=torch.float64, device="cuda").T
A = torch.empty((n, n), dtype=torch.float64, device="cuda").T
tau = torch.empty(n, dtype=torch.float64, device="cuda")
keys = list(grid)
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo):
        h.set_option(k, v)
    ts = []
    for i in range(4):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        if i:
            ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(json.dumps({"n": n, **dict(zip(keys, combo)), "ms": t, "tflops": 4 / 3 * n ** 3 / t / 1e9}), flush=True)
