"""Test-side problem plugins (host Fock builders) for the gscf_b200 solver classes."""
import json
import os

import numpy as np

from gscf_b200 import FocklikeMatrix_i, OrbitalSpace

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "functionality_test.json")))


class Sturmian14Fock(FocklikeMatrix_i):
    """Python twin of gscfmock::FockMatrix (src/gscfmock/FockMatrix.{hh,cc}) used as a *host
    plugin* of the GPU solvers: update({"evec_coefficients": C}) rebuilds F on the CPU."""

    def __init__(self, Z, k_exp, n_alpha, n_beta, guess):
        d = np.load(os.path.join(HERE, "golden", "sturmian14.npz"))
        self.n = 14
        self.t = k_exp * k_exp * d["t_bb_base"]
        self.v0 = -k_exp * Z * d["v0_bb_base"]
        self.s_bb = d["s_bb"].copy()
        self.I4 = (k_exp * d["i_bbbb_base"]).reshape(14, 14, 14, 14)
        self.n_alpha, self.n_beta = n_alpha, n_beta
        self.energies = {}
        self.n_updates = 0
        self._build(np.asarray(guess))

    def _build(self, Cm):
        ca, cb = Cm[:, : self.n_alpha], Cm[:, : self.n_beta]
        pa, pb = ca @ ca.T, cb @ cb.T
        pt = pa + pb
        j = np.einsum("abcd,cd->ab", self.I4, pt)
        ka = np.einsum("acdb,cd->ab", self.I4, pa)
        kb = np.einsum("acdb,cd->ab", self.I4, pb)
        e = {}
        e["energy_kinetic"] = float(np.trace(self.t @ pt))
        e["energy_elec_nuc_attr"] = float(np.trace(self.v0 @ pt))
        e["energy_1e_terms"] = e["energy_kinetic"] + e["energy_elec_nuc_attr"]
        e["energy_coulomb"] = 0.5 * float(np.trace(j @ pt))
        e["energy_exchange"] = -0.5 * float(np.trace(ka @ pa)) - 0.5 * float(np.trace(kb @ pb))
        e["energy_2e_terms"] = e["energy_coulomb"] + e["energy_exchange"]
        e["energy_total"] = e["energy_1e_terms"] + e["energy_2e_terms"]
        self.energies = e
        self.F = self.t + self.v0 + j - ka

    # -- ScfProblemMatrix_i / FocklikeMatrix_i -------------------------------------------------
    def scf_update_key(self):
        return "evec_coefficients"

    def update(self, m):
        if self.scf_update_key() in m:
            self.n_updates += 1
            self._build(np.asarray(m[self.scf_update_key()]))

    def energy_1e_terms(self):
        return self.energies["energy_1e_terms"]

    def energy_2e_terms(self):
        return self.energies["energy_2e_terms"]

    def indices_orbspace(self, osp):
        return {OrbitalSpace.OCC_ALPHA: (0, self.n_alpha), OrbitalSpace.OCC_BETA: (0, self.n_alpha),
                OrbitalSpace.VIRT_ALPHA: (self.n_alpha, self.n), OrbitalSpace.VIRT_BETA: (self.n_beta, self.n)}[osp]

    def n_rows(self):
        return self.n

    def as_dense(self):
        return self.F


def functionality_guess():
    g = np.zeros((14, 4))
    g[0, 0] = -0.9238795325112872
    g[1, 0] = -1.306562964876376
    g[5, 0] = -0.9238795325112864
    g[3, 1] = g[7, 1] = -0.9192110607898044
    g[2, 2] = g[6, 2] = -0.9192110607898044
    g[4, 3] = g[8, 3] = -0.9192110607898044
    return g


def adjust_phase(v, ref):
    return v if np.dot(v, ref) >= 0 else -v
