"""Generalised eigenproblem sweep through the public solver API: a problem matrix that ignores its update, F(D) = H, makes
PlainScf return the eigenpairs of H against S after two iterations.  S ranges from well to badly conditioned; the gate
is what LAPACK dsygvd delivers at the same conditioning (errors grow with cond(S)).

usage: geneig_fuzz.py [n_cases] [first_seed]
"""
import sys

sys.path.insert(0, ".")
import numpy as np
import scipy.linalg as sla

import gscf_b200 as G


class ConstFock(G.FocklikeMatrix_i):
    def __init__(self, H, n_occ):
        self.H, self.n, self.o = H, H.shape[0], n_occ

    def scf_update_key(self):
        return "evec_coefficients"

    def update(self, m):
        pass

    def energy_1e_terms(self):
        return 0.0

    def energy_2e_terms(self):
        return 0.0

    def indices_orbspace(self, osp):
        occ = (0, self.o)
        virt = (self.o, self.n)
        return {G.OrbitalSpace.OCC_ALPHA: occ, G.OrbitalSpace.OCC_BETA: occ, G.OrbitalSpace.VIRT_ALPHA: virt,
                G.OrbitalSpace.VIRT_BETA: virt}[osp]

    def n_rows(self):
        return self.n

    def as_dense(self):
        return self.H


def main(n_cases=None, first=None):
    if n_cases is None:
        n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    if first is None:
        first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for case in range(first, first + n_cases):
        rng = np.random.default_rng(31000 + case)
        n = int(rng.choice([5, 14, 33, 63, 64, 65, 100, 129, 200, 300, 513]))
        logc = float(rng.choice([0.5, 2, 4, 6, 8]))
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        sv = 10.0 ** (-logc * rng.random(n))
        sv[0], sv[-1] = 1.0, 10.0 ** (-logc)
        S = (q * sv) @ q.T
        S = 0.5 * (S + S.T)
        H = rng.standard_normal((n, n))
        H = H + H.T
        o = max(1, n // 8)
        # the error measures ||C_cur - C_prev||-like quantities that vanish for a constant matrix: two iterations
        res = G.PlainScf({"max_error_norm": 1e-3 * 10.0 ** logc, "max_iter": 5}).solve(ConstFock(H, o), S)
        w = np.asarray(res.eigensolution().evalues())
        C = np.asarray(res.eigensolution().evectors())
        wref = sla.eigh(H, S, eigvals_only=True)
        cond = 10.0 ** logc
        scale = np.abs(wref).max()
        tol = 50 * n * 2.2e-16 * cond
        werr = np.abs(w - wref).max() / scale
        orth = np.abs(C.T @ S @ C - np.eye(n)).max()
        resid = np.abs(H @ C - S @ C * w).max() / scale
        ok = werr < tol and orth < tol and resid < tol
        if not ok:
            print(f"FAIL case {case}: n={n} cond=1e{logc:g} werr={werr:.2e} orth={orth:.2e} res={resid:.2e} tol={tol:.2e}",
                  flush=True)
            bad += 1
    print(f"geneig fuzz: {n_cases} cases, failures {bad}")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
