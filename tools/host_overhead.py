"""Host-side cost of one batched call (wall clock per call at a tiny batch, device idle most of the time) and device
time at the per-GPU shard sizes of the strong-scaling run.  usage: python tools/host_overhead.py"""
import sys, time
sys.path.insert(0, "/root/repo")
import torch
from __graft_entry__ import load_package

q = load_package()
h = q.Handle(0)
for n, dt in ((32, torch.float64), (16, torch.float32)):
    A0 = torch.randn((256, n, n), dtype=dt, device="cuda")
    tau = torch.empty((256, n), dtype=dt, device="cuda")
    B = torch.randn((256, 8, n), dtype=dt, device="cuda")
    A = A0.clone()
    for name, fn in (("geqlf", lambda: q.ql_batched_(A, tau, handle=h)),
                     ("lmul", lambda: q.lmul_batched_(A, tau, B, adjoint=True, handle=h))):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(500):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"n={n} {name}: host {1e6 * (t1 - t0) / 500:.1f} us/call issue, {1e6 * (t2 - t0) / 500:.1f} us/call incl. drain")
import itertools
for bv, batch in itertools.product((0, 2), (16384, 32768, 65536, 131072, 262144)):
    h.set_option("bvar", bv)
    A0 = torch.randn((batch, 32, 32), dtype=torch.float64, device="cuda")
    A = A0.clone()
    tau = torch.empty((batch, 32), dtype=torch.float64, device="cuda")
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    ts = []
    for it in range(8):
        A.copy_(A0)
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream()
        h.set_stream(st.cuda_stream) if hasattr(h, "set_stream") else None
        e0.record()
        q.ql_batched_(A, tau, handle=h)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    med = ts[len(ts) // 2]
    print(f"bvar={bv} batch={batch}: {med * 1e3:.1f} us  {batch / med / 1e3:.1f} M/s")
