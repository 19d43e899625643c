"""e2e (host pointers, pinned memory) timing of ql! n=16384: split schedule vs the chunk-joining schedule (run under gpurun)."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = load_package(); h = q.Handle(0)
hp = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
src = torch.randn((n, n), dtype=torch.float64)
Ah = hp.numpy().T
for split, cw in ((1, 0), (0, 0), (1, 6144), (1, 4096), (1, 2048), (1, 0)):
    h.set_option("host_split", split); h.set_option("host_split_chunk", cw)
    ts = []
    for i in range(3):
        hp.copy_(src); torch.cuda.synchronize()
        t0 = time.perf_counter(); q.ql_(Ah, handle=h); ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    print(json.dumps({"n": n, "host_split": split, "chunk": cw, "ms": t * 1e3, "tflops": 4 / 3 * n ** 3 / t / 1e12}), flush=True)
