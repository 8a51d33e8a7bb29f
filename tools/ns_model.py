"""Numerical model (numpy, CPU) of the Newton-Schulz inverse square root used by das_ns_kernel, for
studying accuracy against the conditioning of A = c0 I + Yr^T Y.  It reproduces the kernel's storage
(only block-lower 8x8 tiles are computed, the block-upper half is the mirror) and compares

  old : coupled iteration  M = Z Y, T = sqrt(c)(3I - cM)/2, Z <- T Z, Y <- T Y       (round 1)
  new : product form       T = sqrt(c)(3I - cM)/2, M <- (T M) T, Z <- T Z            (round 2)

against an 80-bit long-double evaluation of A^-1/2 (stable coupled iteration, full storage).
Not part of the product; used to choose the algorithm and its thresholds."""
import numpy as np


def mirror(C, nb=8):
    """keep block-lower tiles (and full diagonal tiles) of C, mirror them into the block-upper half"""
    n = C.shape[0]
    out = C.copy()
    for bi in range(0, n, nb):
        for bj in range(bi + nb, n, nb):
            out[bi:bi + nb, bj:bj + nb] = C[bj:bj + nb, bi:bi + nb].T
    return out


def sprod(X, W):
    return mirror(X @ W)


def truth_invsqrt(A):
    """A^-1/2 in long double: eigen-decomposition in double refined by stable coupled Newton-Schulz"""
    Al = A.astype(np.longdouble)
    lam, V = np.linalg.eigh(A)
    Z = ((V / np.sqrt(lam)) @ V.T).astype(np.longdouble)   # double-accurate start
    n = A.shape[0]
    I = np.eye(n, dtype=np.longdouble)
    for _ in range(4):   # Newton on Z: Z <- Z + (Z R + R Z)/4 is unstable far away but we start at 1e-16*cond
        R = I - Z @ Al @ Z
        # solve the Sylvester correction exactly in the (double) eigenbasis: D_ij = R_ij/(s_i+s_j), Z ~ V s^-1 V^T
        Vl = V.astype(np.longdouble)
        s = np.sqrt(lam).astype(np.longdouble)
        Rt = Vl.T @ R @ Vl
        # (Z+D) A (Z+D) = I, Z = V diag(1/s) V^T  ->  D A Z + Z A D = R  ->  Dt_ij (s_j + s_i) = Rt_ij
        Dt = Rt / (s[:, None] + s[None, :])
        Z = Z + Vl @ Dt @ Vl.T
        Z = (Z + Z.T) / 2
    return Z


def ns_old(A, c0, max_iter=50):
    n = A.shape[0]
    I = np.eye(n)
    G = A - c0 * np.eye(n)
    s = c0 + min(np.abs(G).sum(axis=1).max(), np.sqrt((G * G).sum()) * (1 + 1e-12))
    Y = A / s
    Z = I.copy()
    a, b = c0 / s, 1.0
    it = 0
    first = True
    while True:
        it += 1
        M = Y.copy() if first else sprod(Z, Y)
        res = np.abs(I - M).max()
        last = res < 1e-7 or it >= max_iter
        if (not last) and (not first) and res < 2e-3:
            E = I - M
            E2 = sprod(E, E)
            T = I + 0.5 * E + 0.375 * E2
            if res >= 2e-4:
                T = T + 0.3125 * sprod(E2, E)
            Z = sprod(T, Z)
            return Z / np.sqrt(s), it
        c = 1.0
        if (not last) and (b - a) > 1e-3:
            c = 3.0 / (a + np.sqrt(a * b) + b)
        sc = np.sqrt(c)
        T = 1.5 * sc * I - 0.5 * c * sc * M
        if first:
            Z = T.copy()
            Y = sprod(T, Y)
        else:
            Z = sprod(T, Z)
            if not last:
                Y = sprod(T, Y)
        t = c * a
        a = t * (3 - t) ** 2 / 4
        b = 1.0
        first = False
        if last:
            return Z / np.sqrt(s), (it if res < 1e-7 else -it)


def ns_new(A, c0, max_iter=60, tol_hi=2e-3, variant="tmt"):
    """product form: M is iterated as a function of itself, Z accumulates the T factors"""
    n = A.shape[0]
    I = np.eye(n)
    G = A - c0 * np.eye(n)
    s = c0 + min(np.abs(G).sum(axis=1).max(), np.sqrt((G * G).sum()) * (1 + 1e-12))
    M = A / s
    Z = None
    a, b = c0 / s, 1.0
    it = 0
    while True:
        it += 1
        res = np.abs(I - M).max()
        if res < tol_hi or it >= max_iter:
            E = I - M
            if res < 1e-7:
                T = I + 0.5 * E
            else:
                E2 = sprod(E, E)
                T = I + 0.5 * E + 0.375 * E2
                if res >= 2e-4:
                    T = T + 0.3125 * sprod(E2, E)
            Z = T if Z is None else sprod(T, Z)
            return Z / np.sqrt(s), (it if res < tol_hi else -it)
        c = 1.0
        if (b - a) > 1e-3:
            c = 3.0 / (a + np.sqrt(a * b) + b)
        sc = np.sqrt(c)
        T = 1.5 * sc * I - 0.5 * c * sc * M
        Z = T.copy() if Z is None else sprod(T, Z)
        U = sprod(T, M)
        M = sprod(U, T)
        t = c * a
        a = t * (3 - t) ** 2 / 4
        b = 1.0


def make_A(k, lam_ratio, rank, rng, c0=None, decay="geom"):
    """A = c0 I + G, G PSD of the given rank with ones in its null space, lambda_max(G)/c0 = lam_ratio"""
    c0 = float(k - 1) if c0 is None else c0
    Q, _ = np.linalg.qr(np.column_stack([np.ones(k), rng.standard_normal((k, k - 1))]))
    Q = Q[:, 1:]   # orthonormal basis of the complement of ones
    r = min(rank, k - 1)
    if decay == "geom":
        ev = lam_ratio * c0 * np.logspace(0, -np.log10(max(lam_ratio, 10.0)), r)
    else:
        ev = lam_ratio * c0 * np.ones(r)
    B = Q[:, :r] * np.sqrt(ev)
    G = B @ B.T
    G = (G + G.T) / 2
    return c0 * np.eye(k) + G, c0


def pad(A, n):
    k = A.shape[0]
    P = np.eye(n)
    P[:k, :k] = A
    return P


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    print(f"{'k':>4} {'ratio':>8} {'rank':>5} | {'old err':>10} {'it':>3} | {'new err':>10} {'it':>3} | resid_new")
    for k in (20, 50, 100):
        n = 8 * ((k + 2 + 7) // 8)
        if (n // 8) % 2 == 0:
            n += 8
        for ratio in (1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7):
            for rank in (5, k - 1):
                A, c0 = make_A(k, ratio, rank, rng)
                Zt = truth_invsqrt(A)
                nrm = float(np.abs(Zt).max())
                Ap = pad(A / 1.0, n)
                # padding rows are the identity *after* scaling in the kernel; emulate: scale then pad
                Zo, ito = ns_old_padded = ns_old(A, c0)
                Zn, itn = ns_new(A, c0)
                eo = float(np.abs(Zo - Zt).max() / nrm)
                en = float(np.abs(Zn - Zt).max() / nrm)
                Rn = np.abs(np.eye(k) - Zn @ A @ Zn).max()
                print(f"{k:4d} {ratio:8.0e} {rank:5d} | {eo:10.2e} {ito:3d} | {en:10.2e} {itn:3d} | {Rn:.2e}")
