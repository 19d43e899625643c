import sys, numpy as np
sys.path.insert(0, "/root/repo")
from __graft_entry__ import load_package
q = load_package()
h0, h1 = q.Handle(0), q.Handle(1)
rng = np.random.default_rng(1)
A = rng.standard_normal((100001, 32, 32))
F1, t1 = q.ql_batched_(A.copy(), handle=h0)
F2, t2 = q.ql_batched_mg_(A.copy(), handles=[h0, h1])
print("mg equal:", np.array_equal(F1, F2) and np.array_equal(t1, t2))
