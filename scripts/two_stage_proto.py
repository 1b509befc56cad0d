"""Round-2 groundwork: numpy prototype of the two-stage tridiagonalisation (dense -> band -> tridiagonal) with both
back-transformations, to fix the algorithm, its data flow and its operation counts before the CUDA version
(DESIGN.md section 5 has the cost estimate for B200; nothing here is on the product path).

Stage 1 (sy2sb), block column k of width b, m = n - (k+1) b rows below the band:
    QR of the m x b panel            -> V (unit lower trapezoidal), T (b x b), R            [panel kernel: 16-CTA cluster]
    X = A22 V T ; W = X - V (T^T (V^T X)) / 2 ; A22 -= V W^T + W V^T                          [DMMA GEMMs, 2 n^3 total]
Stage 2 (sb2st), sweep j: a length-b reflector annihilates column j below the sub-diagonal, the bulge it creates is
    chased down the band by ~(n - j) / b further reflectors, each a two-sided update of a (2b x b)-ish window.
    Sweep j+1 may start as soon as sweep j is two windows ahead: ~3 n dependent steps on the critical path.
Back-transformation: Z = Q1 (Q2 Z_T); Q2's reflectors are applied in reverse order of creation, grouped in the CUDA version
    into diamond-shaped blocks for GEMM efficiency; Q1 is the usual compact-WY product (2 n^3).

usage: two_stage_proto.py [n] [b]
"""
import sys

import numpy as np


def house(x):
    """LAPACK dlarfg: (v, tau, beta) with (I - tau v v^T) x = beta e1, v[0] = 1."""
    alpha, xn = x[0], np.linalg.norm(x[1:])
    v = np.zeros_like(x)
    v[0] = 1.0
    if xn == 0.0:
        return v, 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    v[1:] = x[1:] / (alpha - beta)
    return v, (beta - alpha) / beta, beta


def panel_qr(P):
    """Householder QR of an m x b panel: V (m x b), T (b x b upper), R (b x b upper) with P = (I - V T V^T) [R; 0]."""
    P = P.copy()
    m, b = P.shape
    V = np.zeros((m, b))
    taus = np.zeros(b)
    for c in range(min(b, m)):
        v, tau, beta = house(P[c:, c])
        V[c:, c] = v
        taus[c] = tau
        P[c:, c + 1:] -= tau * np.outer(v, v @ P[c:, c + 1:])
        P[c, c] = beta
        P[c + 1:, c] = 0.0
    T = np.zeros((b, b))
    for c in range(b):  # dlarft, forward columnwise
        T[c, c] = taus[c]
        if c:
            T[:c, c] = -taus[c] * (T[:c, :c] @ (V[:, :c].T @ V[:, c]))
    return V, T, np.triu(P[:b, :])


def sy2sb(A, b):
    """Dense symmetric -> symmetric band of half-bandwidth b.  Returns (band matrix, [(row offset, V, T)])."""
    A = A.copy()
    n = A.shape[0]
    refl = []
    flops = 0.0
    for k in range(0, n - b - 1, b):
        r0 = k + b
        m = n - r0
        V, T, R = panel_qr(A[r0:, k:k + b])
        A[r0:, k:k + b] = 0.0
        A[r0:r0 + R.shape[0], k:k + b] = R[:min(b, m), :]
        A[k:k + b, r0:] = A[r0:, k:k + b].T
        A22 = A[r0:, r0:]
        X = A22 @ V @ T
        W = X - 0.5 * V @ (T.T @ (V.T @ X))
        A[r0:, r0:] = A22 - V @ W.T - W @ V.T
        flops += 2.0 * m * m * b + 4.0 * m * m * b  # symmetric product + full-storage rank-2b update
        refl.append((r0, V, T))
    return A, refl, flops


def sb2st(B, b):
    """Band -> tridiagonal by bulge chasing.  Returns (d, e, [(row offset, v, tau)] in order of application, steps)."""
    B = B.copy()
    n = B.shape[0]
    refl = []
    for j in range(n - 2):
        col, r0 = j, j + 1
        while r0 < n - 1:
            r1 = min(r0 + b, n)
            x = B[r0:r1, col]
            if len(x) < 2 or not np.any(x[1:]):
                break
            v, tau, beta = house(x.copy())
            # two-sided application on the window; the prototype touches whole rows / columns, the kernel only the
            # (r1 - r0) x ~3b window that is non-zero
            B[r0:r1, :] -= tau * np.outer(v, v @ B[r0:r1, :])
            B[:, r0:r1] -= tau * np.outer(B[:, r0:r1] @ v, v)
            refl.append((r0, v, tau))
            col, r0 = r0, r0 + b
    return np.diag(B).copy(), np.diag(B, -1).copy(), refl


def back_transform(Zt, refl1, refl2):
    Z = Zt.copy()
    for r0, v, tau in reversed(refl2):  # Q2 = H_1 H_2 ... H_m
        blk = Z[r0:r0 + len(v), :]
        blk -= tau * np.outer(v, v @ blk)
    for r0, V, T in reversed(refl1):    # Q1 = (I - V_1 T_1 V_1^T) (I - V_2 T_2 V_2^T) ...
        blk = Z[r0:, :]
        blk -= V @ (T @ (V.T @ blk))
    return Z


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    rng = np.random.default_rng(0)
    A = rng.standard_normal((n, n))
    A = A + A.T
    B, refl1, flops1 = sy2sb(A, b)
    assert np.abs(np.tril(B, -b - 1)).max() < 1e-12 * n, "not a band matrix"
    d, e, refl2 = sb2st(B, b)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w, Zt = np.linalg.eigh(T)
    Z = back_transform(Zt, refl1, refl2)
    wref = np.linalg.eigvalsh(A)
    print(f"n={n} b={b}: eigenvalue error {np.abs(w - wref).max():.2e}  orthogonality {np.abs(Z.T @ Z - np.eye(n)).max():.2e}  "
          f"residual {np.abs(A @ Z - Z * w).max():.2e}")
    print(f"stage 1: {len(refl1)} panels, {flops1 / 1e9:.3f} GF in GEMMs (2 n^3 = {2 * n ** 3 / 1e9:.3f}); "
          f"stage 2: {len(refl2)} reflectors of length <= {b} (n^2 / b = {n * n // b}), "
          f"critical path ~ {3 * n} dependent window updates")


if __name__ == "__main__":
    main()
