import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np
import gscf_b200 as G
from gscf_b200 import lib as L

def run(n, o, aux, m, method="diis", P=1, eig=0, reps=2, n_beta=None):
    fock = G.DeviceDFFock(n, o, n_beta, n_aux=aux, seeds=list(range(P)))
    for r in range(reps):
        t0 = time.time()
        eng, st = G.solve_device_problem(fock, method, {"diis_n_prev_steps": m}, eig_method=eig, phase_timing=True)
        wall = time.time() - t0
        its = eng.n_iter()
        scf, fk = eng.timings()
        ph = eng.phase_timings()
        e1, e2 = eng.energies()
        eng.close()
        print(f"n={n} P={P} {method} st={st} iters={its[:4]}.. sum={its.sum()} wall={wall*1e3:.1f}ms scf={scf:.2f}ms fock={fk:.2f}ms "
              f"it/s={its.max()/scf*1e3:.1f} E0={e1[0]+e2[0]:.10f}")
        print("   phases:", {k: round(v, 3) for k, v in ph.items() if k != 'counts'})
        print("   counts:", ph["counts"])
    fock.close()

if __name__ == "__main__":
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    L.check(L.lib.gscf_b200_measure_peaks(C.byref(a), C.byref(b), C.byref(c)))
    print(f"peaks: DMMA {a.value:.2f} TF/s, DFMA {b.value:.2f} TF/s, copy {c.value:.1f} GB/s")
    which = sys.argv[1:] or ["c2"]
    if "c2" in which: run(1024, 128, 64, 8)
    if "c3" in which: run(4096, 512, 64, 10, reps=1)
    if "c4" in which: run(128, 16, 8, 4, P=4096)
    if "c5full" in which: run(512, 64, 32, 4, P=256, n_beta=60, reps=1)
    if "c5" in which: run(512, 64, 32, 4, P=16, n_beta=60)
    if "mid" in which: run(512, 64, 32, 8)
    if "c4dc" in which: run(128, 16, 8, 4, P=4096, eig=2, reps=2)
