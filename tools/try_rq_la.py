"""rq 65536 x 1024 (configs[3]) with the look-ahead threshold lowered (run under gpurun)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
mr, nr = 65536, 1024
Ar = torch.randn((nr, mr), dtype=torch.float64, device="cuda").T
Arw = torch.empty((nr, mr), dtype=torch.float64, device="cuda").T
tr_ = torch.empty(nr, dtype=torch.float64, device="cuda")
for la_min, nb in ((4096, 0), (1024, 0), (1024, 128), (1024, 192), (1024, 256), (512, 128)):
    h.set_option("la_min", la_min); h.set_option("nb", nb)
    ts = []
    for i in range(5):
        Arw.T.copy_(Ar.T)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(); h.call("qlb200_dgerqf_dev", mr, nr, Arw.data_ptr(), Arw.stride(1), tr_.data_ptr()); b1.record()
        torch.cuda.synchronize()
        if i: ts.append(b0.elapsed_time(b1))
    fl = 2.0 * nr * nr * (mr - nr / 3.0)
    print("la_min", la_min, "nb", nb, "ms", min(ts), "TFLOP/s", fl / min(ts) / 1e9, flush=True)
