"""Randomised sweep of gscf_b200_dgemm against numpy: shapes around the tile edges, all transposes, even/odd leading
dimensions (TMA needs 16-byte aligned rows, cp.async 8-byte), batches with strides, beta = 0 with NaN-filled C.

usage: gemm_fuzz.py [n_cases] [first_seed]
"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from gscf_b200 import lib as L

EDGE = [1, 2, 3, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 130, 191, 192, 200, 255, 256, 257, 300, 511,
        512, 513, 640, 1000, 1024, 1025]


def colmajor(a, ld):
    m, n = a.shape
    buf = np.zeros((n, ld))
    buf[:, :m] = a.T
    return buf.ravel()


def main(n_cases=None, first=None):
    if n_cases is None:
        n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    if first is None:
        first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for case in range(first, first + n_cases):
        rng = np.random.default_rng(77000 + case)
        m, n = int(rng.choice(EDGE)), int(rng.choice(EDGE))
        k = int(rng.choice([0] + EDGE))
        if m * n * max(k, 1) > 3e8:
            k = int(rng.choice([16, 33, 64]))
        ta, tb = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        batch = int(rng.choice([1, 1, 2, 5]))
        if batch > 1 and m * n > 300 * 300:
            batch = 1
        alpha = float(rng.choice([1.0, -1.0, 1.5]))
        beta = float(rng.choice([0.0, 1.0, -0.25]))
        ra, ca = (k, m) if ta else (m, k)
        rb, cb = (n, k) if tb else (k, n)
        lda = max(1, ra) + int(rng.choice([0, 1, 2, 3]))
        ldb = max(1, rb) + int(rng.choice([0, 1, 2, 3]))
        ldc = m + int(rng.choice([0, 1, 2]))
        A = rng.standard_normal((batch, ra, ca))
        B = rng.standard_normal((batch, rb, cb))
        Cm = rng.standard_normal((batch, m, n))
        sA, sB, sC = lda * max(1, ca) + int(rng.choice([0, 4])), ldb * max(1, cb) + int(rng.choice([0, 6])), ldc * n
        bufA = np.zeros(batch * sA)
        bufB = np.zeros(batch * sB)
        bufC = np.zeros(batch * sC)
        for b in range(batch):
            if ra * ca:
                bufA[b * sA: b * sA + lda * ca] = colmajor(A[b], lda)
            if rb * cb:
                bufB[b * sB: b * sB + ldb * cb] = colmajor(B[b], ldb)
            cm = colmajor(Cm[b], ldc)
            if beta == 0.0:
                cm = np.full_like(cm, np.nan)  # beta = 0 must not read C (BLAS semantics)
            bufC[b * sC: b * sC + ldc * n] = cm
        dA, dB, dC = (torch.from_numpy(x).cuda() for x in (bufA, bufB, bufC))
        desc = f"case {case}: ta={ta} tb={tb} m={m} n={n} k={k} lda={lda} ldb={ldb} ldc={ldc} batch={batch} a={alpha} b={beta}"
        try:
            L.check(L.lib.gscf_b200_dgemm(ta, tb, m, n, k, alpha, dA.data_ptr(), lda, sA, dB.data_ptr(), ldb, sB, beta,
                                          dC.data_ptr(), ldc, sC, batch, None))
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            print("FAIL (exception)", desc, repr(ex), flush=True)
            bad += 1
            continue
        got = dC.cpu().numpy()
        for b in range(batch):
            g = got[b * sC: b * sC + ldc * n].reshape(n, ldc)[:, :m].T
            opA = A[b].T if ta else A[b]
            opB = B[b].T if tb else B[b]
            ref = alpha * (opA @ opB) + (beta * Cm[b] if beta != 0.0 else 0.0)
            tol = 1e-13 * max(1, k) * (1 + np.abs(ref).max())
            err = np.abs(g - ref).max() if g.size else 0.0
            if not (err < tol):
                print("FAIL", desc, "member", b, "err", err, flush=True)
                bad += 1
                break
    print(f"gemm fuzz: {n_cases} cases, failures {bad}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
