"""Generate tests/golden/*.npz -- golden input/output vectors for the QL hot path.

The reference (Julia) cannot run in this image (no julia binary).  For
Float64/Float32 its factorization arithmetic IS the LAPACK routine ?geqlf (src/ql.jl:72) and
?gerqf (src/rq.jl:401); the identical reference-LAPACK routines are exported by SciPy's bundled
OpenBLAS, so the golden outputs are produced by calling exactly those (oracle/lapack.py).  The
Q-apply / solve / UL goldens are produced by the line-faithful C restatement of the reference's
scalar Julia loops (oracle/ql_oracle_impl.h) after that restatement has been pinned against
LAPACK (?ormql, ?getrf) in tests/test_oracle.py.

Run from the repo root:  python tests/golden/make_golden.py
Inputs are seeded numpy (PCG64, seed 1234321 mirroring test/test_qr.jl:9) -- Julia's Xoshiro stream
is not reproducible outside Julia, so literal reference vectors are not portable; the reference
tests pin results relationally (flip-QR identity, sign test), which tests/ re-check as well.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cport, lapack  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(1234321)
    out = {}
    # single matrices: the shapes of test/test_ql.jl:5-33,151-181 plus blocked-path sizes
    for name, (m, n), dt in [
        ("sq10", (10, 10), np.float64), ("wide10x12", (10, 12), np.float64), ("tall12x10", (12, 10), np.float64),
        ("wide3x5", (3, 5), np.float64), ("sq100", (100, 100), np.float64), ("wide100x103", (100, 103), np.float64),
        ("sq200", (200, 200), np.float64), ("tall300x130", (300, 130), np.float64), ("sq64_f32", (64, 64), np.float32),
    ]:
        A = (rng.standard_normal((m, n)) / 2).astype(dt)
        F, tau = lapack.geqlf(A)
        out[f"{name}_A"], out[f"{name}_F"], out[f"{name}_tau"] = A, F, tau
        if m == n:
            B = rng.standard_normal((m, 3)).astype(dt)
            out[f"{name}_B"] = B
            out[f"{name}_QtB"] = cport.lmul(F, tau, B, adjoint=True)
            out[f"{name}_QB"] = cport.lmul(F, tau, B, adjoint=False)
            out[f"{name}_X"] = cport.ldiv(F, tau, B)
    # RQ (src/rq.jl:401 -> gerqf)
    for name, (m, n) in [("rq_sq10", (10, 10)), ("rq_wide8x13", (8, 13)), ("rq_tall40x12", (40, 12))]:
        A = rng.standard_normal((m, n)) / 2
        F, tau = lapack.gerqf(A)
        out[f"{name}_A"], out[f"{name}_F"], out[f"{name}_tau"] = A, F, tau
    # UL (src/ul.jl:148-196)
    for name, n in [("ul10", 10), ("ul33", 33)]:
        A = rng.standard_normal((n, n))
        F, ipiv, info = cport.ulfact(A)
        out[f"{name}_A"], out[f"{name}_F"], out[f"{name}_ipiv"] = A, F, ipiv
    # batched: (batch, n, n) C-order == Julia Array{T,3}(n,n,batch) memory
    for name, n, batch, dt, nrhs in [("b32_f64", 32, 16, np.float64, 8), ("b16_f32", 16, 32, np.float32, 8),
                                      ("b8_f64", 8, 8, np.float64, 2)]:
        A = rng.standard_normal((batch, n, n)).astype(dt)
        F = A.copy()
        tau = lapack.geqlf_loop(F)
        B = rng.standard_normal((batch, nrhs, n)).astype(dt)   # batch column-major n x nrhs
        QtB = np.stack([cport.lmul(F[i].T, tau[i], B[i].T, adjoint=True).T for i in range(batch)])
        X = np.stack([cport.ldiv(F[i].T, tau[i], B[i].T).T for i in range(batch)])
        out[f"{name}_A"], out[f"{name}_F"], out[f"{name}_tau"] = A, F, tau
        out[f"{name}_B"], out[f"{name}_QtB"], out[f"{name}_X"] = B, QtB, X
    # Issue 7304 sign test (test/test_ql.jl:120-124): known answer Q = -A
    A = np.array([[-np.sqrt(.5), -np.sqrt(.5)], [-np.sqrt(.5), np.sqrt(.5)]])
    F, tau = lapack.geqlf(A)
    out["i7304_A"], out["i7304_F"], out["i7304_tau"] = A, F, tau
    np.savez_compressed(os.path.join(HERE, "ql_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ql_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
