"""Quick batched QL timing over bvar variants (run under gpurun): bb.py n dtype batch bvar,bvar,..."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import cport  # noqa: E402
n = int(sys.argv[1]); dt = torch.float64 if sys.argv[2] == "f64" else torch.float32
batch = int(sys.argv[3]); bvars = [int(x) for x in sys.argv[4].split(",")]
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1234321)
A0 = torch.randn(batch, n, n, dtype=dt, device="cuda", generator=g); A = A0.clone()
tau = torch.empty(batch, n, dtype=dt, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
sub = 256
Fo, tauo = cport.geql2_batched(A0[:sub].cpu().numpy())
for bv in bvars:
    h.set_option("bvar", bv)
    ts = []
    for i in range(9):
        A.copy_(A0); flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); q.ql_batched_(A, tau, handle=h); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    err = float(np.abs(A[:sub].cpu().numpy() - Fo).max()); terr = float(np.abs(tau[:sub].cpu().numpy() - tauo).max())
    es = A.element_size()
    print(json.dumps({"n": n, "dtype": sys.argv[2], "bvar": bv, "ms": round(ms, 4), "Mmat_s": round(batch / ms / 1e3, 1),
                      "gbs": round(batch * (2 * n * n + n) * es / ms / 1e6, 1), "err": err, "tau_err": terr}))
