"""e2e (host pointers, pinned memory) timing of ql! n=16384: the three host schedules of capi.cu fact_host (run under gpurun)."""
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
for split, cw, ramp, jc in ((2, 1024, 4, 2), (1, 1024, 4, 2), (0, 1024, 4, 2)):
    h.set_option("host_split", split); h.set_option("chunk_cols", cw); h.set_option("la_join_ramp", ramp); h.set_option("la_join_chunks", jc)
    ts = []
    first = None
    for i in range(5):
        hp.copy_(src); torch.cuda.synchronize()
        t0 = time.perf_counter(); F = q.ql_(Ah, handle=h); ts.append(time.perf_counter() - t0)
        if i == 0: first = (hp[:64].clone(), hp[-64:].clone(), F.tau.copy())
        else: assert torch.equal(first[0], hp[:64]) and torch.equal(first[1], hp[-64:]) and (first[2] == F.tau).all(), "replay differs"
    print([round(x * 1e3, 1) for x in ts], h.launches)
    t = min(ts[1:])
    print(json.dumps({"n": n, "host_split": split, "chunk": cw, "ramp": ramp, "jc": jc, "ms": t * 1e3, "tflops": 4 / 3 * n ** 3 / t / 1e12}), flush=True)
