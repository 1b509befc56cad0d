"""Stress sweep of gscf_b200_syev over orders and special matrices (run on the GPU box; prints failures).

usage: eig_stress.py [quick]
"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from gscf_b200 import lib as L


def colmajor(a, ld):
    n = a.shape[0]
    out = np.zeros((n, ld))
    out[:, :n] = a.T
    return out.ravel()


def kinds(n, rng):
    a = rng.standard_normal((n, n))
    a = a + a.T
    yield "random", a
    yield "zero", np.zeros((n, n))
    yield "identity", np.eye(n)
    yield "diag", np.diag(np.linspace(-1.0, 1.0, n))
    yield "rank1", np.outer(np.ones(n), np.ones(n))
    yield "tiny", a * 1e-150
    yield "huge", a * 1e150
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.repeat(np.arange(1, n // 4 + 2), 4)[:n].astype(float)
    b = (q * ev) @ q.T
    yield "clustered", 0.5 * (b + b.T)
    w = np.diag(np.abs(np.arange(n) - (n - 1) / 2.0)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    yield "wilkinson", w
    # glued Wilkinson blocks: many eigenvalues agree to rounding (hard for D&C deflation / orthogonality)
    g = w.copy()
    for i in range(20, n - 1, 21):
        g[i, i + 1] = g[i + 1, i] = 1e-8
    yield "glued", g
    blk = np.zeros((n, n))
    h = n // 2
    blk[:h, :h] = a[:h, :h]
    blk[h:, h:] = a[:n - h, :n - h]
    yield "blockdiag", blk
    yield "graded", a * np.outer(10.0 ** -np.linspace(0, 12, n), 10.0 ** -np.linspace(0, 12, n))


def run(method, n, mats):
    batch = len(mats)
    ld = L.lib.gscf_b200_padded_ld(n)
    buf = np.concatenate([colmajor(a, ld) for a in mats])
    dA = torch.from_numpy(buf).cuda()
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.full((batch,), 7, dtype=torch.int32, device="cuda")
    wsz = L.lib.gscf_b200_syev_worksize(n, batch, method)
    work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_syev(method, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld, ld * n,
                                 batch, work.data_ptr(), wsz, info.data_ptr(), None))
    torch.cuda.synchronize()
    return info.cpu().numpy(), dw.cpu().numpy().reshape(batch, n), dZ.cpu().numpy().reshape(batch, n, ld)[:, :, :n]


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    orders = [3, 4, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 130, 144, 145, 150, 151, 200, 255,
              256, 257, 300, 400, 511, 512, 513, 592, 593, 700, 1000, 1023, 1024, 1025, 1100]
    if not quick:
        orders += [1500, 1663, 1664, 1665, 1792, 1793, 1800, 2049]
    bad = 0
    t0 = time.time()
    for n in orders:
        rng = np.random.default_rng(n)
        names, mats = zip(*kinds(n, rng))
        for method in ([L.EIG_AUTO, L.EIG_TRIDIAG_DC] if n <= 64 else [L.EIG_AUTO]):
            info, W, Z = run(method, n, list(mats))
            for b, (name, a) in enumerate(zip(names, mats)):
                wref = np.linalg.eigvalsh(a)
                scale = max(np.abs(wref).max(), np.finfo(float).tiny)
                z = Z[b].T
                werr = np.abs(W[b] - wref).max() / scale
                orth = np.abs(z.T @ z - np.eye(n)).max()
                res = np.abs(a @ z - z * W[b]).max() / scale
                tol = 2e-13 * n
                ok = info[b] == 0 and werr < tol and orth < tol and res < tol
                if not ok:
                    bad += 1
                    print(f"FAIL n={n} method={method} {name}: info={info[b]} werr={werr:.2e} orth={orth:.2e} res={res:.2e}",
                          flush=True)
        print(f"n={n} done ({time.time() - t0:.0f} s)", flush=True)
    print("failures:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
