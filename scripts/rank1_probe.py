"""Probe: orthogonality of the eigenvectors of the all-ones matrix (n-1 fold eigenvalue 0) and of its tridiagonal form."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gscf_b200 import lib as L
from scripts.eig_stress import run


def tridiag(a):
    a = a.copy(); n = a.shape[0]; d = np.zeros(n); e = np.zeros(n - 1)
    for j in range(n - 1):
        x = a[j + 1:, j].copy()
        alpha = x[0]; xn = np.linalg.norm(x[1:])
        if xn == 0:
            tau = 0.0; beta = alpha; v = np.zeros_like(x); v[0] = 1
        else:
            beta = -np.copysign(np.hypot(alpha, xn), alpha); tau = (beta - alpha) / beta; v = x / (alpha - beta); v[0] = 1
        e[j] = beta; d[j] = a[j, j]
        A22 = a[j + 1:, j + 1:]
        y = tau * (A22 @ v); w = y - 0.5 * tau * (y @ v) * v
        a[j + 1:, j + 1:] = A22 - np.outer(v, w) - np.outer(w, v)
    d[n - 1] = a[n - 1, n - 1]
    return d, e


for n in [int(x) for x in sys.argv[1:]] or [128, 255, 256, 257]:
    a = np.ones((n, n))
    d, e = tridiag(a)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    for name, m in [("dense ones", a), ("its tridiagonal", T)]:
        info, W, Z = run(L.EIG_AUTO, n, [m])
        z = Z[0].T
        G = z.T @ z - np.eye(n)
        i, j = np.unravel_index(np.abs(G).argmax(), G.shape)
        res = np.abs(m @ z - z * W[0]).max()
        print(f"n={n} {name}: orth {np.abs(G).max():.2e} at ({i},{j}) w=({W[0][i]:.3e},{W[0][j]:.3e}) res {res:.2e} "
              f"bad pairs {(np.abs(G) > 1e-10).sum() // 2} |col norms-1| {np.abs(np.diag(G)).max():.2e}", flush=True)
