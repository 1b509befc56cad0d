import sys, numpy as np
sys.path.insert(0, '.')
import gscf_b200 as G
from gscf_b200 import lib as L
from oracle import truncated_opt_damp_scf
from oracle.synthetic import SyntheticDFProblem
n,o,aux=128,16,16
fock = G.DeviceDFFock(n, o, n_aux=aux, seeds=[0])
eng, st = G.solve_device_problem(fock, "toda", {})
en, er, cf = eng.trace(0)
p = SyntheticDFProblem(n, o, n_aux=aux, seed=0)
C0 = p.core_guess()[None]
F0, e1, e2 = p.fock(C0)
ref = truncated_opt_damp_scf(p, F0, e1, e2)
print('gpu its', eng.n_iter(), 'ref', ref.n_iter)
for i in range(max(len(en), len(ref.trace_energy))):
    a = (en[i], er[i], float(cf[i][0])) if i < len(en) else None
    b = (ref.trace_energy[i], ref.trace_error[i], ref.trace_coeff[i]) if i < len(ref.trace_energy) else None
    print(i+1, a, b)
