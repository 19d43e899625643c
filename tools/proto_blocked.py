"""NumPy model of the device-side blocked QL driver (csrc/blocked.cu): same panel/strip
decomposition, same workspaces (Vw, T, W), each device kernel replaced by the dense operation it
performs.  Used on the CPU to pin the index arithmetic of the driver before it runs on a GPU
(tests/test_proto_blocked.py compares it with LAPACK ?geqlf).  Not part of the product."""
import numpy as np


def larfg(alpha, x):
    xnorm = np.linalg.norm(x)
    if xnorm == 0.0:
        return alpha, x, 0.0
    beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
    tau = (beta - alpha) / beta
    return beta, x / (alpha - beta), tau


def strip_kernel(A, ps, s0, s1, tau, tau_off, Vw, vcol0, p):
    """Unblocked QL of A[0:ps, s0:s1] (pivot of column c at row ps - (s1 - c)); writes the cleaned
    reflectors into Vw[0:p, vcol0:vcol0+w] and returns the strip's T (w x w, lower)."""
    w = s1 - s0
    G = np.zeros((w, w))
    for c in range(s1 - 1, s0 - 1, -1):
        mu = ps - (s1 - c)
        if mu < 0:
            continue
        x = A[:mu, c].copy()
        # one fused reduction: d_j = x . A[0:mu, j] for every column j of the strip
        d = x @ A[:mu, s0:s1]
        rowv = A[mu, s0:s1].copy()
        beta, v, tk = larfg(A[mu, c], x)
        scale = 0.0 if tk == 0 else 1.0 / (A[mu, c] - beta)
        A[:mu, c] = v
        A[mu, c] = beta
        tau[tau_off + (c - s0)] = tk
        for j in range(s0, c):
            s = rowv[j - s0] + d[j - s0] * scale
            t = tk * s
            A[mu, j] -= t
            A[:mu, j] -= t * scale * x
        for j in range(c + 1, s1):
            G[j - s0, c - s0] = rowv[j - s0] + scale * d[j - s0]     # v_j . v_c
    T = np.zeros((w, w))
    for j in range(w - 1, -1, -1):
        tj = tau[tau_off + j]
        T[j, j] = tj
        T[j + 1:, j] = -tj * (T[j + 1:, j + 1:] @ G[j + 1:, j])
    for c in range(s0, s1):
        mu = ps - (s1 - c)
        col = vcol0 + (c - s0)
        Vw[:p, col] = 0.0
        if mu >= 0:
            Vw[:mu, col] = A[:mu, c]
            Vw[mu, col] = 1.0
    return T


def larft_from_gram(G, tau):
    ib = G.shape[0]
    T = np.zeros((ib, ib))
    for j in range(ib - 1, -1, -1):
        T[j, j] = tau[j]
        T[j + 1:, j] = -tau[j] * (T[j + 1:, j + 1:] @ G[j + 1:, j])
    return T


def geqlf_blocked(A, nb=128, sb=16):
    A = np.array(A, dtype=np.float64, order="F", copy=True)
    m, n = A.shape
    k = min(m, n)
    tau = np.zeros(k)
    Vw = np.zeros((m, nb))
    c1 = n
    while c1 > n - k:
        c0 = max(n - k, c1 - nb)
        ib = c1 - c0
        p = m - n + c1
        s1 = c1
        while s1 > c0:
            s0 = max(c0, s1 - sb)
            ps = m - n + s1
            Ts = strip_kernel(A, ps, s0, s1, tau, s0 - (n - k), Vw, s0 - c0, p)
            if s0 > c0:      # apply the strip's block reflector to the panel columns left of it
                Vs = Vw[:ps, s0 - c0:s1 - c0]
                C = A[:ps, c0:s0]
                W = C.T @ Vs
                W = W @ Ts
                C -= Vs @ W.T
            s1 = s0
        if c0 > 0:
            V = Vw[:p, :ib]
            T = larft_from_gram(V.T @ V, tau[c0 - (n - k):c1 - (n - k)])
            C = A[:p, :c0]
            W = C.T @ V
            W = W @ T
            C -= V @ W.T
        c1 = c0
    return A, tau


def ormql_blocked(side, trans, F, tau, C, nb=128):
    """C <- op(Q) C (side L) or C op(Q) (side R) from the packed factors (LAPACK ?ormql order)."""
    F = np.asarray(F)
    C = np.array(C, dtype=np.float64, order="F", copy=True)
    mA, nA = F.shape
    k = min(mA, nA)
    nq = mA
    blocks = []
    i = 0
    while i < k:
        ib = min(nb, k - i)
        blocks.append((i, ib))
        i += ib
    left, notran = side == "L", trans == "N"
    # Q = H_k ... H_1: Q C applies H_1 first (ascending); Q' C descending; C Q ascending means ... see below
    if (left and notran) or (not left and not notran):
        order = blocks
    else:
        order = blocks[::-1]
    for (i, ib) in order:
        c0 = nA - k + i
        p = nq - k + i + ib
        V = np.zeros((p, ib))
        for j in range(ib):
            mu = nq - k + i + j
            V[:mu, j] = F[:mu, c0 + j]
            V[mu, j] = 1.0
        T = larft_from_gram(V.T @ V, tau[i:i + ib])
        # left: H C = C - V T (V'C) -> W2 = W T' for Q, W T for Q'; right: C H = C - (C V) T V' -> W T / W T'
        Tm = (T.T if notran else T) if left else (T if notran else T.T)
        if left:
            W = C[:p, :].T @ V
            W = W @ Tm
            C[:p, :] -= V @ W.T
        else:
            W = C[:, :p] @ V
            W = W @ Tm
            C[:, :p] -= W @ V.T
    return C
