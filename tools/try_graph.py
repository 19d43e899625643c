"""Does replaying ql! n=16384 as a CUDA graph beat plain launches? (run under gpurun)"""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = load_package(); h = q.Handle(0)
A0 = torch.randn((n, n), dtype=torch.float64, device="cuda").T
A = torch.empty((n, n), dtype=torch.float64, device="cuda").T
tau = torch.empty(n, dtype=torch.float64, device="cuda")
s = torch.cuda.Stream()
def run():
    h.call("qlb200_dgeqlf_dev", n, n, A.data_ptr(), A.stride(1), tau.data_ptr())
with torch.cuda.stream(s):
    h.set_stream(s.cuda_stream)
    for _ in range(2):
        A.T.copy_(A0.T); run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        A.T.copy_(A0.T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("plain ms", min(ts))
    ref_tau = tau.clone()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    run()
ts = []
for _ in range(3):
    A.T.copy_(A0.T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("graph ms", min(ts), "same result:", bool(torch.equal(tau, ref_tau)))
