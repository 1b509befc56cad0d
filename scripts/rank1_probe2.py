import sys
sys.path.insert(0, ".")
import numpy as np, torch
from gscf_b200 import lib as L
from scripts.eig_stress import run
for n in [int(x) for x in sys.argv[2:]] or [129, 256]:
    batch = int(sys.argv[1])
    a = np.ones((n, n))
    info, W, Z = run(L.EIG_AUTO, n, [a] * batch)
    for b in sorted(set([0, batch - 1])):
        z = Z[b].T
        G = z.T @ z
        nr = np.sqrt(np.diag(G))
        badc = np.where(np.abs(nr - 1) > 1e-8)[0]
        off = np.abs(G - np.diag(np.diag(G))).max()
        print(f"n={n} batch={batch} b={b}: max|colnorm-1| {np.abs(nr - 1).max():.2e} ncols bad {len(badc)} first bad {badc[:8]} "
              f"norms {nr[badc[:4]]} w there {W[b][badc[:4]]} offdiag {off:.2e} max|Z| {np.abs(z).max():.2e}", flush=True)
        if len(badc):
            c = badc[0]
            print("   bad column top entries:", np.sort(np.abs(z[:, c]))[-4:], " w range", W[b][:3], W[b][-2:])
