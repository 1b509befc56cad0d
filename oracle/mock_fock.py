"""gscfmock restated (TEST ORACLE): the Sturmian-14 closed-shell mock Fock plugin.

Follows ``src/gscfmock/FockMatrix.cc:25-206`` and
``src/gscfmock/Integrals/IntegralsSturmian14.hh:37-41``.  The integral tables
are test data: ``tests/golden/make_golden.py`` parses them out of the reference
checkout into ``tests/golden/sturmian14.npz``; nothing here reads
``/root/reference`` at run time.
"""
from __future__ import annotations

import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class IntegralsSturmian14:
    """``IntegralsSturmian14(Z, k)``: T*k^2, V0*(-k Z), ERI*k  (IntegralsSturmian14.hh:37-41)."""

    def __init__(self, Z: float, k_exp: float, path: str | None = None):
        d = np.load(path or os.path.join(_GOLDEN, "sturmian14.npz"))
        self.nbas = 14
        self.t_bb = (k_exp * k_exp) * d["t_bb_base"]
        self.v0_bb = (-k_exp * Z) * d["v0_bb_base"]
        self.s_bb = d["s_bb"].copy()
        self.i_bbbb = k_exp * d["i_bbbb_base"]  # (n*n, n*n): I[(a n + b), (c n + d)]


class MockFockProblem:
    """Closed-shell mock Fock matrix: F = T + V0 + J(Pa+Pb) - K(Pa).

    ``fock(C)`` = ``FockMatrix::update`` -> ``build_fock_matrix_from_coefficient``
    (FockMatrix.cc:192-206) -> ``build_fock_matrix_from_density`` (:126-190).
    """

    def __init__(self, idata: IntegralsSturmian14, n_alpha: int, n_beta: int):
        # The reference asserts closed shell (FockMatrix.cc:177).  n_alpha != n_beta is an extension used to test the
        # unrestricted (block-diagonal) branch of the solvers (ScfBase.hh:300-312) - PARITY UNPINNED: the same J / K /
        # energy formulas (they are written per spin in FockMatrix.cc:157-167), F^s = T + V0 + J(Pa + Pb) - K(P^s).
        self.unrestricted = n_alpha != n_beta
        self.idata = idata
        self.n_bas = idata.nbas
        self.n_alpha, self.n_beta = n_alpha, n_beta
        self.S = idata.s_bb
        n = self.n_bas
        self._I4 = idata.i_bbbb.reshape(n, n, n, n)  # [a, b, c, d]
        self.last_energies = {}

    def fock(self, C: np.ndarray):
        C = np.asarray(C)
        if C.ndim == 2:
            C = C[None]
        ca = C[0][:, : self.n_alpha]
        cb = C[1 if (self.unrestricted and C.shape[0] > 1) else 0][:, : self.n_beta]
        pa = ca @ ca.T  # FockMatrix.cc:202-203
        pb = cb @ cb.T
        pt = pa + pb
        I4 = self._I4
        # J_ab = sum_cd P_cd I[(ab),(cd)]      (FockMatrix.cc:57-91)
        j = np.einsum("abcd,cd->ab", I4, pt)
        # K_ab = sum_cd P_cd I[(ac),(db)]      (FockMatrix.cc:93-124)
        ka = np.einsum("acdb,cd->ab", I4, pa)
        kb = np.einsum("acdb,cd->ab", I4, pb)
        idata = self.idata
        e_kin = float(np.trace(idata.t_bb @ pt))  # :157-160
        e_v0 = float(np.trace(idata.v0_bb @ pt))
        e1 = e_kin + e_v0
        e_coul = 0.5 * float(np.trace(j @ pt))  # :163-167
        e_xchg = -0.5 * float(np.trace(ka @ pa)) - 0.5 * float(np.trace(kb @ pb))
        e2 = e_coul + e_xchg
        F = idata.t_bb + idata.v0_bb + j - ka  # :179-182
        self.last_energies = dict(
            energy_kinetic=e_kin,
            energy_elec_nuc_attr=e_v0,
            energy_1e_terms=e1,
            energy_coulomb=e_coul,
            energy_exchange=e_xchg,
            energy_2e_terms=e2,
            energy_total=e1 + e2,
        )
        if self.unrestricted:
            return np.stack([F, idata.t_bb + idata.v0_bb + j - kb]), e1, e2
        return F[None].copy(), e1, e2


def loewdin_guess(S: np.ndarray) -> np.ndarray:
    """``examples/hartree_fock/loewdin_guess.cc:26-46``: eigenvectors of S
    (ascending eigenvalue), each scaled by 1/sqrt(eigenvalue)."""
    w, v = np.linalg.eigh(S)
    return v / np.sqrt(w)[None, :]


def functionality_test_guess() -> np.ndarray:
    """The hand-typed 14x4 guess of ``tests/FunctionalityTest.cc:51-57``."""
    g = np.zeros((14, 4))
    g[0, 0] = -0.9238795325112872
    g[1, 0] = -1.306562964876376
    g[5, 0] = -0.9238795325112864
    g[3, 1] = g[7, 1] = -0.9192110607898044
    g[2, 2] = g[6, 2] = -0.9192110607898044
    g[4, 3] = g[8, 3] = -0.9192110607898044
    return g
