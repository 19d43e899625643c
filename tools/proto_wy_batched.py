"""NumPy model of the shared-memory batched QL kernel (batched_wy.cu): 8-column panels factored with the ROTATING column
scheme (pivot column always at panel index 7, finished reflectors re-enter at index 0, one uniform step body), the
compact-WY factor S (M = H_0 ... H_7 = I - V S V') accumulated from the Gram terms the step already reduces, and the
trailing update C <- C - V S (V' C).  Checked against LAPACK dgeqlf (same sign conventions)."""
import numpy as np


def ql_wy(A, pad=0):
    A = A.copy()
    n = A.shape[0]
    assert n % 8 == 0
    tau = np.zeros(n)
    for p in range(n // 8 - 1, -1, -1):
        if 8 * p + 7 < pad:
            break
        rows = 8 * p + 8
        P = [A[:rows, 8 * p + i].copy() for i in range(8)]     # panel columns (rows below the panel are final L)
        Srot = np.zeros((8, 8))                                  # Srot[i, c] = rotating column store of lane c
        for r in range(7, -1, -1):
            R = 8 * p + r
            active = R > pad
            x = P[7]
            xm = np.where(np.arange(rows) < R, x, 0.0)
            d = np.array([xm @ P[i] for i in range(8)])
            aR = np.array([P[i][R] for i in range(8)])
            ssq, alpha = d[7], aR[7]
            if active:
                nrm = np.sqrt(alpha * alpha + ssq)
                beta = -np.copysign(nrm, alpha)
                tk = (beta - alpha) / beta
                scale = 1.0 / (alpha - beta)
            else:
                beta, tk, scale = alpha, 0.0, 0.0
            tau[R] = tk
            w = d * scale + aR                                  # i < 7: v_r' a_i (live) or v_r' v_c (finished)
            live = np.array([i + r >= 7 for i in range(8)])
            s = np.where(live, tk * w, 0.0)
            acc = np.array([sum(w[i] * Srot[i, c] for i in range(7)) for c in range(8)])
            Snew = np.where(np.arange(8) == r, tk, -tk * acc)
            Srot[1:8] = Srot[0:7].copy()
            Srot[0] = Snew
            v = xm * scale
            vf = v.copy()
            vf[R] = 1.0
            old7 = P[7].copy()
            for i in range(6, -1, -1):
                P[i + 1] = P[i] - s[i] * vf
            P0 = np.where(np.arange(rows) < R, v, old7)
            P0[R] = beta
            P[0] = P0
        for i in range(8):
            A[:rows, 8 * p + i] = P[i]
        S = Srot.copy()                                          # S[i, c] natural order
        if p >= 1:
            V = np.array(P).T.copy()                             # rows x 8, packed
            for c in range(8):                                   # clean: unit pivot, zeros below
                V[8 * p + c, c] = 1.0
                V[8 * p + c + 1:, c] = 0.0
            C = A[:rows, :8 * p]
            W = C.T @ V
            W2 = -W @ S.T
            A[:rows, :8 * p] = C + (W2 @ V.T).T
    return A, tau


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import lapack
    rng = np.random.default_rng(5)
    for n, pad in [(8, 0), (16, 0), (32, 0), (64, 0), (48, 0), (32, 8), (32, 5), (64, 31), (64, 17), (40, 7)]:
        nr = n - pad
        Ar = rng.standard_normal((nr, nr))
        A = np.zeros((n, n))
        A[pad:, pad:] = Ar
        F, tau = ql_wy(A, pad)
        Fo, tauo = lapack.geqlf(np.asfortranarray(Ar))
        print(n, pad, np.abs(F[pad:, pad:] - Fo).max(), np.abs(tau[pad:] - tauo).max())
