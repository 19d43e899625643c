"""Device-side timings of the 8(f) rows: qr via flips at n = 8192/16384, tall least squares 65536 x 1024 with 8 right-hand sides.
usage: python tools/time_qr_lstsq.py"""
import sys
sys.path.insert(0, "/root/repo")
import torch
from __graft_entry__ import load_package

q = load_package()
h = q.Handle(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
h.set_stream(st.cuda_stream)


def timed(fn, reset, iters=3):
    ts = []
    for _ in range(iters + 1):
        reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])


for n in (8192, 16384):
    A0 = torch.randn((n, n), dtype=torch.float64, device="cuda").T
    A = (A0.T.clone()).T
    tau = None
    t_ql = timed(lambda: q.ql_(A, handle=h), lambda: A.T.copy_(A0.T))
    t_qr = timed(lambda: q.qr_(A, handle=h), lambda: A.T.copy_(A0.T))
    fl = 4 / 3 * n ** 3
    print(f"n={n}: ql! {t_ql:.2f} ms ({fl / t_ql / 1e9:.2f} TFLOP/s)   qr (flip, ql!, flip) {t_qr:.2f} ms ({fl / t_qr / 1e9:.2f} TFLOP/s)")
    del A0, A
m, n, nrhs = 65536, 1024, 8
A0 = torch.randn((n, m), dtype=torch.float64, device="cuda").T
A = (A0.T.clone()).T
F = q.ql_(A, handle=h)
B0 = torch.randn((nrhs, m), dtype=torch.float64, device="cuda").T
B = (B0.T.clone()).T
t = timed(lambda: q.ldiv_(F, B), lambda: B.T.copy_(B0.T))
X = B[m - n:]
r = A0 @ X - B0
print(f"lstsq {m}x{n}, nrhs={nrhs}: {t:.2f} ms after ql!; ||A'(Ax-b)||/(||A||^2 ||x||) = "
      f"{float(torch.linalg.norm(A0.T @ r) / (torch.linalg.norm(A0) ** 2 * torch.linalg.norm(X))):.2e}")
