"""Randomised SCF parity sweep: engine (GPU) against the oracle on small synthetic problems.

Random n, occupations (restricted / unrestricted), auxiliary count, solver, DIIS depth, ensemble size and eigensolver.
Gates: same iteration count, |dE| <= 1e-10, max |d eps| <= 1e-9 for every ensemble member; members the oracle cannot
converge within max_iter must end with ERR_MAX_ITER on the GPU as well.  Cases whose decisive error norm lies within 0.5 %
of the threshold are skipped (iteration parity is then a property of the rounding).

usage: scf_fuzz.py [n_cases] [first_seed] [large]
"""
import sys

sys.path.insert(0, ".")
import numpy as np

import gscf_b200 as G
from gscf_b200 import lib as L
from oracle import ScfFailed, plain_scf, pulay_diis_scf, truncated_opt_damp_scf
from oracle.synthetic import SyntheticDFProblem


def oracle_run(n, na, nb, aux, which, seed, **kw):
    p = SyntheticDFProblem(n, na, None if na == nb else nb, n_aux=aux, seed=seed)
    C0 = p.core_guess()
    C0 = np.stack([C0, C0]) if p.unrestricted else C0[None]
    F0, e1, e2 = p.fock(C0)
    fn = {"plain": plain_scf, "diis": pulay_diis_scf, "toda": truncated_opt_damp_scf}[which]
    return fn(p, F0, e1, e2, **kw)


LARGE = [200, 257, 320, 400, 520]


def main(n_cases=None, first=None, large=False):
    if n_cases is None:
        n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    if first is None:
        first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = skipped = checked = 0
    for case in range(first, first + n_cases):
        rng = np.random.default_rng(1000 + case)
        n = int(rng.choice(LARGE if large else [6, 9, 14, 20, 31, 32, 33, 48, 64, 65, 80, 96, 130, 150]))
        na = int(rng.integers(1, max(2, n // 4)))
        nb = na if rng.random() < 0.6 else int(rng.integers(1, na + 1))
        aux = int(rng.integers(2, 10))
        which = str(rng.choice(["plain", "diis", "diis", "toda"]))
        m = int(rng.integers(1, 9))
        P = int(rng.choice([1, 1, 2, 5]))
        eig = int(rng.choice([L.EIG_AUTO, L.EIG_JACOBI, L.EIG_TRIDIAG_DC])) if n <= 150 else L.EIG_AUTO
        if large:
            P = min(P, 2)
        thr = 1e-5 if which == "toda" else float(rng.choice([5e-7, 1e-6, 1e-8]))
        max_iter = 40
        params = {"max_error_norm": thr, "max_iter": max_iter}
        kw = dict(max_error_norm=thr, max_iter=max_iter)
        if which == "diis":
            params["diis_n_prev_steps"] = m
            kw["n_prev_steps"] = m
        keig = n
        if rng.random() < 0.3:  # partial spectrum (ScfBase.hh:87): same iterates, only the lowest k pairs are returned
            keig = int(rng.integers(max(na, nb), n + 1))
            params["n_eigenpairs"] = keig
        seeds = [5000 + 10 * case + i for i in range(P)]
        desc = f"case {case}: n={n} na={na} nb={nb} aux={aux} {which} m={m} P={P} eig={eig} thr={thr:g} k={keig}"
        fock = G.DeviceDFFock(n, na, nb, n_aux=aux, seeds=seeds)
        try:
            eng, st = G.solve_device_problem(fock, which, params, eig_method=eig)
        except Exception as ex:  # noqa: BLE001
            print("FAIL (exception)", desc, repr(ex), flush=True)
            bad += 1
            fock.close()
            continue
        try:
            its, conv, status = eng.n_iter(), eng.converged(), eng.status()
            e1, e2 = eng.energies()
            w = eng.orben()
        finally:
            eng.close()
            fock.close()
        for i, s in enumerate(seeds):
            try:
                ref = oracle_run(n, na, nb, aux, which, s, **kw)
            except ScfFailed as ex:
                if ex.kind == "max_iter":
                    if conv[i] != 0 or status[i] != L.ERR_MAX_ITER or its[i] != max_iter:
                        print("FAIL (oracle hits max_iter)", desc, "member", i, its[i], conv[i], status[i], flush=True)
                        bad += 1
                    else:
                        checked += 1
                else:
                    skipped += 1
                continue
            errs = np.asarray(ref.trace_error)
            if np.any(np.abs(errs / thr - 1.0) < 5e-3):
                skipped += 1
                continue
            ok = (conv[i] == 1 and its[i] == ref.n_iter and abs(e1[i] + e2[i] - ref.energy_total) < 1e-10)
            tol_w = 1e-6 if which == "toda" else 1e-9
            ok = ok and np.abs(w[i][:, :keig] - ref.orben[:, :keig]).max() < tol_w
            if not ok:
                print("FAIL", desc, "member", i, "iters", its[i], ref.n_iter, "conv", conv[i], "status", status[i],
                      "dE", abs(e1[i] + e2[i] - ref.energy_total), "dw", np.abs(w[i][:, :keig] - ref.orben[:, :keig]).max(), flush=True)
                bad += 1
            else:
                checked += 1
    print(f"checked {checked} members, skipped {skipped}, failures {bad}")
    return bad, checked


if __name__ == "__main__":
    sys.exit(1 if main(large=len(sys.argv) > 3 and sys.argv[3] == "large")[0] else 0)
