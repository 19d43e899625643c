"""Small pass over every kernel family (for compute-sanitizer): tiny sizes, host-pointer entry points."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
q = load_package(); h = q.Handle(0)
rng = np.random.default_rng(0)
for (m, n) in [(130, 130), (70, 45), (45, 70), (17, 1)]:
    A = np.asfortranarray(rng.standard_normal((m, n)))
    F = q.ql(A, handle=h)
    B = np.asfortranarray(rng.standard_normal((m, 3)))
    q.lmul_(F.Q, B.copy(order="F")); q.lmul_(F.Q.T, B.copy(order="F"))
    C = np.asfortranarray(rng.standard_normal((4, m)))
    q.rmul_(C.copy(order="F"), F.Q); q.rmul_(C.copy(order="F"), F.Q.T)
    if m == n:
        F.solve(B)
A = np.asfortranarray(rng.standard_normal((40, 12))); Fr, tr = q.rq_(A.copy(order="F"), handle=h)
Q = q.RQPackedQ(Fr, tr, handle=h); q.lmul_(Q, np.asfortranarray(rng.standard_normal((12, 2))))
for n in (5, 33, 100):
    A = np.asfortranarray(rng.standard_normal((n, n))); U = q.ul(A, handle=h); U.solve(np.asfortranarray(rng.standard_normal((n, 2))))
for dt in (np.float64, np.float32):
    for n in (8, 16, 32, 12, 48):
        A3 = rng.standard_normal((37, n, n)).astype(dt)
        F3, t3 = q.ql_batched_(A3.copy(), handle=h)
        if n in (8, 16, 32):
            for nr in (1, 2, 3, 8, 11):
                B3 = rng.standard_normal((37, nr, n)).astype(dt)
                q.lmul_batched_(F3, t3, B3.copy(), adjoint=True, handle=h)
                q.lmul_batched_(F3, t3, B3.copy(), adjoint=False, handle=h)
                q.ldiv_batched_(F3, t3, B3.copy(), handle=h)
h.sync()
print("sanity pass done; launches", h.launches)
