"""e2e (host pointers, pinned memory) timing of ql! n=16384 for several join_step values (run under gpurun)."""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
n = 16384
q = load_package(); h = q.Handle(0)
hp = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
src = torch.randn((n, n), dtype=torch.float64)
Ah = hp.numpy().T
for spec in sys.argv[1].split(","):
    cw, js = [int(x) for x in spec.split(":")]
    if js == 0:
        h.set_option("stream_in", 0)
    else:
        h.set_option("stream_in", 1); h.set_option("join_step", js); h.set_option("chunk_cols", cw)
    ts = []
    for i in range(3):
        hp.copy_(src); torch.cuda.synchronize()
        t0 = time.perf_counter(); q.ql_(Ah, handle=h); ts.append(time.perf_counter() - t0)
    t = min(ts[1:])
    print(json.dumps({"chunk_cols": cw, "join_step": js, "ms": t * 1e3, "tflops": 4 / 3 * n ** 3 / t / 1e12}), flush=True)
