"""Small invocations of the round-2 kernels for compute-sanitizer (run under gpurun):
  compute-sanitizer --tool memcheck python tools/sanity_round2.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

q = load_package()
h = q.Handle(0)
rng = np.random.default_rng(0)
for n, dt in [(5, np.float64), (12, np.float64), (24, np.float64), (31, np.float32), (40, np.float64), (47, np.float64), (48, np.float32),
              (56, np.float64), (64, np.float64), (63, np.float64), (64, np.float32)]:
    A = rng.standard_normal((37, n, n)).astype(dt)
    F, tau = q.ql_batched_(A.copy(), handle=h)
    B = rng.standard_normal((37, 3, n)).astype(dt)
    q.lmul_batched_(F, tau, B.copy(), adjoint=True, handle=h)
    q.ldiv_batched_(F, tau, B.copy(), handle=h)
    print("batched", n, dt.__name__, float(np.abs(F).max()))
for bv in (6, 7):
    h.set_option("bvar", bv)
    A = rng.standard_normal((20, 48, 48))
    q.ql_batched_(A, handle=h)
h.set_option("bvar", 0)
q.ql_apply_batched_(rng.standard_normal((50, 32, 32)), rng.standard_normal((50, 8, 32)), op="S", handle=h)
for shape in [(70, 70), (200, 130), (130, 200)]:
    A = np.asfortranarray(rng.standard_normal(shape) + 1j * rng.standard_normal(shape))
    F = q.ql(A, handle=h)
    q.lmul_(F.Q.T, np.asfortranarray(rng.standard_normal((shape[0], 3)) + 0j))
    q.rmul_(np.asfortranarray(rng.standard_normal((3, shape[0])) + 0j), F.Q)
    if shape[0] == shape[1]:
        F.solve(np.asfortranarray(rng.standard_normal((shape[0], 2)) + 0j))
    print("complex", shape)
A = np.asfortranarray(rng.standard_normal((300, 300)).astype(np.float32))
F = q.ql(A, handle=h)
F.solve(np.asfortranarray(rng.standard_normal((300, 2)).astype(np.float32)))
A = np.asfortranarray(rng.standard_normal((2200, 2200)))
q.ul(A, handle=h)
print("done")
