"""Special matrices at an order that takes the deferred-update (blocked) kernel and the hand-over: batch of all kinds at
once (no hand-over: one matrix per launch) and one by one (batch 1: hand-over to the shared-memory kernel)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from gscf_b200 import lib as L
from scripts.eig_stress import kinds, run

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2601
names, mats = zip(*kinds(n, np.random.default_rng(n)))
refs = [np.linalg.eigvalsh(a) for a in mats]
bad = 0
for mode in ("batch", "single"):
    groups = [list(range(len(mats)))] if mode == "batch" else [[i] for i in range(len(mats))]
    for grp in groups:
        info, W, Z = run(L.EIG_AUTO, n, [mats[i] for i in grp])
        for b, i in enumerate(grp):
            a, wref = mats[i], refs[i]
            scale = max(np.abs(wref).max(), np.finfo(float).tiny)
            z = Z[b].T
            werr = np.abs(W[b] - wref).max() / scale
            orth = np.abs(z.T @ z - np.eye(n)).max()
            res = np.abs(a @ z - z * W[b]).max() / scale
            ok = info[b] == 0 and max(werr, orth, res) < 2e-13 * n
            if not ok:
                bad += 1
            print(f"{'ok  ' if ok else 'FAIL'} n={n} {mode} {names[i]}: info={info[b]} werr={werr:.2e} orth={orth:.2e} res={res:.2e}", flush=True)
print("failures:", bad)
