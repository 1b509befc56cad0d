"""Special matrices in LARGE batches: the batched tridiagonalisation modes (one CTA per matrix in registers / shared
memory, clusters of 4 CTAs with the matrix in L2, several launches per batch) see every kind at every batch position."""
import sys
sys.path.insert(0, ".")
import numpy as np
from gscf_b200 import lib as L
from scripts.eig_stress import kinds, run

bad = 0
for n, reps in [(128, 25), (150, 9), (512, 4), (592, 2)]:
    names, mats = zip(*kinds(n, np.random.default_rng(n)))
    refs = [np.linalg.eigvalsh(a) for a in mats]
    order = np.random.default_rng(1).permutation(len(mats) * reps)
    batch = [mats[i % len(mats)] for i in order]
    info, W, Z = run(L.EIG_AUTO, n, batch)
    for b, i in enumerate(order):
        k = i % len(mats)
        a, wref = mats[k], refs[k]
        scale = max(np.abs(wref).max(), np.finfo(float).tiny)
        z = Z[b].T
        werr = np.abs(W[b] - wref).max() / scale
        orth = np.abs(z.T @ z - np.eye(n)).max()
        res = np.abs(a @ z - z * W[b]).max() / scale
        if not (info[b] == 0 and max(werr, orth, res) < 2e-13 * n):
            bad += 1
            print(f"FAIL n={n} batch pos {b} {names[k]}: info={info[b]} werr={werr:.2e} orth={orth:.2e} res={res:.2e}", flush=True)
    print(f"n={n} batch={len(batch)} done", flush=True)
print("failures:", bad)
