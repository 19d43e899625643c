"""Host-pointer batched calls (pinned NumPy arrays) with and without the chunk pipeline.
usage: python tools/try_batched_e2e.py"""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from __graft_entry__ import load_package

q = load_package()
h = q.Handle(0)
batch, n, nrhs = 262144, 32, 8
src = torch.randn((batch, n, n), dtype=torch.float64)
Ah = torch.empty((batch, n, n), dtype=torch.float64).pin_memory()
th = torch.empty((batch, n), dtype=torch.float64).pin_memory()
Bh = torch.randn((batch, nrhs, n), dtype=torch.float64).pin_memory()
A, tau, B = Ah.numpy(), th.numpy(), Bh.numpy()
ref = None
for mb in (0, 8, 32, 128):
    h.set_option("batch_chunk_mb", mb)
    ts, tl = [], []
    for it in range(4):
        Ah.copy_(src)
        t0 = time.perf_counter()
        q.ql_batched_(A, tau, handle=h)
        t1 = time.perf_counter()
        q.lmul_batched_(A, tau, B, adjoint=True, handle=h)
        t2 = time.perf_counter()
        ts.append(t1 - t0)
        tl.append(t2 - t1)
    if ref is None:
        ref = (A.copy(), tau.copy())
    same = np.array_equal(ref[0], A) and np.array_equal(ref[1], tau)
    print(f"chunk {mb:4d} MB: ql {min(ts) * 1e3:7.2f} ms ({batch / min(ts) / 1e6:6.2f} M/s)  lmul {min(tl) * 1e3:7.2f} ms  same bits: {same}")
