"""TODA at the default threshold: GPU (Jacobi / tridiagonal eigensolver) against the oracle, iteration by iteration."""
import sys

import numpy as np

sys.path.insert(0, ".")
import gscf_b200 as G  # noqa: E402
from gscf_b200 import lib as L  # noqa: E402
from oracle import truncated_opt_damp_scf  # noqa: E402
from oracle.synthetic import SyntheticDFProblem  # noqa: E402

CASES = [(48, 6, 6, 1), (64, 8, 8, 0), (64, 8, 8, 2), (32, 4, 4, 0), (56, 7, 7, 3)]
for n, o, aux, seed in CASES:
    p = SyntheticDFProblem(n, o, n_aux=aux, seed=seed)
    F0, e1, e2 = p.fock(p.core_guess()[None])
    ref = truncated_opt_damp_scf(p, F0, e1, e2, max_iter=200)
    res = {}
    for name, eig in (("auto", L.EIG_AUTO), ("dc", L.EIG_TRIDIAG_DC)):
        fock = G.DeviceDFFock(n, o, n_aux=aux, seeds=[seed])
        eng, st = G.solve_device_problem(fock, "toda", {"max_iter": 200}, eig_method=eig)
        res[name] = (int(eng.n_iter()[0]), eng.trace(0))
        eng.close()
        fock.close()
    print(f"case n={n} seed={seed}: oracle {ref.n_iter}  gpu auto {res['auto'][0]}  gpu dc {res['dc'][0]}")
    for i in range(max(ref.n_iter, res['auto'][0], res['dc'][0])):
        row = [f"{i + 1:3d}"]
        row.append(f"ref lam={ref.trace_coeff[i]:.6f} err={ref.trace_error[i]:.3e}" if i < ref.n_iter else " " * 34)
        for name in ("auto", "dc"):
            k, (en, er, cf) = res[name]
            row.append(f"{name} lam={float(cf[i][0]):.6f} err={er[i]:.3e} dE={en[i] - ref.trace_energy[min(i, ref.n_iter - 1)]:+.2e}"
                       if i < k else "")
        print("  ".join(row))
