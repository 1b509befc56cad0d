"""Restatement of gscf's three SCF drivers on numpy/scipy-LAPACK (TEST ORACLE).

Every function cites the reference lines it follows (paths relative to the
reference checkout).  Conventions: matrices are numpy float64; a "blocked"
quantity has shape ``(nblk, n, n)`` with ``nblk == 1`` for restricted problems
and ``nblk == 2`` (alpha, beta) for unrestricted ones.  Coefficient columns are
orbitals, sorted by ascending orbital energy.

The problem plugin (``src/gscf/ScfProblemMatrix_i.hh:32-44`` +
``src/gscf/FocklikeMatrix_i.hh:53-117``) is any object with

    n_bas, n_alpha, n_beta          sizes; restricted iff a single block
    S                               (n, n) overlap
    fock(C) -> (F, E1e, E2e)        F(C) as (nblk, n, n) and the two energies
                                    (``update({{key, C}})`` + ``energy_*_terms``)
"""
from __future__ import annotations

import dataclasses
from collections import deque
from typing import List, Optional

import numpy as np
import scipy.linalg as sla


class ScfFailed(RuntimeError):
    """Mirror of lazyten::SolverException (max-iter, DIIS step, eigensolver)."""

    def __init__(self, kind: str, msg: str, result: "Optional[ScfResult]" = None):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind
        self.result = result


@dataclasses.dataclass
class ScfResult:
    n_iter: int
    converged: bool
    last_error_norm: float
    orben: np.ndarray  # (nblk, n)
    C: np.ndarray  # (nblk, n, n)
    F: np.ndarray  # (nblk, n, n) problem matrix built from C
    energy_1e: float
    energy_2e: float
    # per-iteration records (observability)
    trace_energy: List[float]
    trace_error: List[float]
    trace_coeff: List[object]  # DIIS coefficient vectors or TODA damping

    @property
    def energy_total(self) -> float:
        return self.energy_1e + self.energy_2e

    def density(self, n_alpha: int, n_beta: int) -> np.ndarray:
        """D_sigma = C_occ C_occ^T per block (src/gscfmock/FockMatrix.cc:198-203)."""
        occ = [n_alpha, n_beta]
        return np.stack(
            [self.C[b][:, : occ[b]] @ self.C[b][:, : occ[b]].T for b in range(self.C.shape[0])]
        )


# --------------------------------------------------------------------------------------
# building blocks of ScfBase
# --------------------------------------------------------------------------------------
_EIG_DRIVER = "gvd"


class eig_driver:
    """Context manager selecting the LAPACK route of ``generalised_eigh`` (test infrastructure: used to show how far two
    LAPACK builds of the *same* algorithm drift apart, tests/test_toda_sensitivity_cpu.py).  ``gvd`` (default, divide &
    conquer ``dsygvd``), ``gv`` (QR iteration ``dsygv``), ``loewdin`` (X = S^-1/2, ``dsyevd`` of X F X)."""

    def __init__(self, name: str):
        assert name in ("gvd", "gv", "loewdin"), name
        self.name = name

    def __enter__(self):
        global _EIG_DRIVER
        self.saved, _EIG_DRIVER = _EIG_DRIVER, self.name
        return self

    def __exit__(self, *exc):
        global _EIG_DRIVER
        _EIG_DRIVER = self.saved
        return False


def generalised_eigh(F: np.ndarray, S: np.ndarray):
    """F C = S C eps, all pairs, ascending, C^T S C = I.

    ``src/gscf/ScfBase.hh:291-339`` delegates to lazyten's EigensystemSolver
    (dense: armadillo/LAPACK generalised symmetric solver); restated with
    LAPACK ``dsygvd`` (other routes only through ``eig_driver``).
    """
    if _EIG_DRIVER == "loewdin":
        ws, U = np.linalg.eigh(S)
        X = (U * ws**-0.5) @ U.T
        w, Z = np.linalg.eigh(X @ F @ X)
        return w, X @ Z
    w, v = sla.eigh(F, S, driver=_EIG_DRIVER, check_finite=False)
    return w, v


def _eig_blocks(Fdiag: np.ndarray, S: np.ndarray):
    ws, vs = [], []
    for b in range(Fdiag.shape[0]):
        w, v = generalised_eigh(Fdiag[b], S)
        ws.append(w)
        vs.append(v)
    return np.stack(ws), np.stack(vs)


def pulay_error(F: np.ndarray, C: np.ndarray, S: np.ndarray, n_alpha: int, n_beta: int):
    """Pulay error, ``src/gscfmock/pulay_error.cc:25-44``.

    Restricted (one block): e = (2 S C_o)(F C_o)^T - (F C_o)(2 S C_o)^T.
    Unrestricted (PARITY UNPINNED - the reference asserts closed shell at
    pulay_error.cc:30): one such block per spin without the closed-shell factor
    2, e_s = (S C_so)(F_s C_so)^T - (F_s C_so)(S C_so)^T.
    """
    nblk = F.shape[0]
    occ = [n_alpha, n_beta]
    out = np.empty_like(F)
    for b in range(nblk):
        co = C[b][:, : occ[b]]
        fac = 2.0 if nblk == 1 else 1.0
        sc = fac * (S @ co)
        fc = F[b] @ co
        out[b] = sc @ fc.T - fc @ sc.T
    return out


def _norm_frobenius(e: np.ndarray) -> float:
    return float(np.sqrt(np.sum(e * e)))


def _check_convergence(n_iter, last_err, max_error_norm, max_iter):
    """lazyten IterativeWrapper::convergence_reached [upstream-memory]:
    ``is_converged`` (``src/gscf/ScfBase.hh:123-126``) first, then the max_iter
    assert.  ``last_error_norm`` starts as NaN (``ScfStateBase.hh:171``) so the
    first test is false."""
    if last_err < max_error_norm:
        return True
    if n_iter >= max_iter:
        raise ScfFailed("max_iter", f"maximum number of iterations ({max_iter}) reached")
    return False


# --------------------------------------------------------------------------------------
# PlainScf  (src/gscf/PlainScf.hh:121-147)
# --------------------------------------------------------------------------------------
def plain_scf(problem, F0, e1_0, e2_0, max_error_norm=5e-7, max_iter=100) -> ScfResult:
    S = problem.S
    F, e1, e2 = np.array(F0, copy=True), e1_0, e2_0
    n_iter, last_err = 0, float("nan")
    orben = C = None
    tE, tR, tC = [], [], []
    while not _check_convergence(n_iter, last_err, max_error_norm, max_iter):
        n_iter += 1  # start_iteration_step
        orben, C = _eig_blocks(F, S)  # PlainScf.hh:129-130
        F, e1, e2 = problem.fock(C)  # PlainScf.hh:140 -> ScfBase.hh:341-370
        last_err = _norm_frobenius(  # PlainScf.hh:143 -> ScfBase.hh:220-222
            pulay_error(F, C, S, problem.n_alpha, problem.n_beta)
        )
        tE.append(e1 + e2)
        tR.append(last_err)
        tC.append(None)
    return ScfResult(n_iter, True, last_err, orben, C, F, e1, e2, tE, tR, tC)


# --------------------------------------------------------------------------------------
# PulayDiisScf  (src/gscf/PulayDiisScf.hh:283-463)
# --------------------------------------------------------------------------------------
def diis_linear_system_matrix(error_overlaps) -> np.ndarray:
    """``PulayDiisScf.hh:326-356``: B(i, i-j) = B(i-j, i) = overlaps[i][j]; border -1; corner 0."""
    k = len(error_overlaps)
    B = np.zeros((k + 1, k + 1))
    for i, ov in enumerate(error_overlaps):
        B[i, i] = ov[0]
        for j in range(1, i + 1):
            B[i, i - j] = B[i - j, i] = ov[j]
        B[k, i] = B[i, k] = -1.0
    B[k, k] = 0.0
    return B


def _diis_coefficients(error_overlaps) -> np.ndarray:
    """``PulayDiisScf.hh:358-393``; linear solve = LAPACK dgesv semantics."""
    k = len(error_overlaps)
    if k == 0:
        return np.zeros(0)
    if k == 1:
        return np.ones(1)
    B = diis_linear_system_matrix(error_overlaps)
    rhs = np.zeros(k + 1)
    rhs[k] = -1.0
    try:
        x = np.linalg.solve(B, rhs)
    except np.linalg.LinAlgError as exc:  # lazyten::SolverException -> ExcDiisStepFailed
        raise ScfFailed("diis_step", str(exc))
    if not np.all(np.isfinite(x)):
        raise ScfFailed("diis_step", "non-finite DIIS coefficients")
    return x[:k]


def pulay_diis_scf(
    problem, F0, e1_0, e2_0, n_prev_steps=4, max_error_norm=5e-7, max_iter=100
) -> ScfResult:
    S = problem.S
    F, e1, e2 = np.array(F0, copy=True), e1_0, e2_0
    prev_F = deque(maxlen=n_prev_steps)  # PulayDiisScf.hh:56, resize :271-277
    errors = deque(maxlen=n_prev_steps)  # :60
    overlaps = deque(maxlen=n_prev_steps)  # :81
    n_iter, last_err = 0, float("nan")
    orben = C = None
    tE, tR, tC = [], [], []
    while not _check_convergence(n_iter, last_err, max_error_norm, max_iter):
        n_iter += 1
        c = _diis_coefficients(overlaps)  # :295
        if len(c) == 0:  # :402-406
            Fdiag = F
        else:  # :414-428, oldest -> newest
            assert len(c) == len(prev_F)
            Fdiag = c[0] * prev_F[0]
            for ci, Fi in zip(c[1:], list(prev_F)[1:]):
                Fdiag = Fdiag + ci * Fi
        orben, C = _eig_blocks(Fdiag, S)  # :299
        F, e1, e2 = problem.fock(C)  # :314
        prev_F.append(F)  # :315
        e = pulay_error(F, C, S, problem.n_alpha, problem.n_beta)
        errors.append(e)  # :318
        last_err = _norm_frobenius(e)  # :319 -> :435-438
        overlaps.append(  # :320 -> :440-463, newest first
            [float(np.sum(e * ej)) for ej in reversed(errors)]
        )
        tE.append(e1 + e2)
        tR.append(last_err)
        tC.append(np.array(c, copy=True))
    return ScfResult(n_iter, True, last_err, orben, C, F, e1, e2, tE, tR, tC)


# --------------------------------------------------------------------------------------
# TruncatedOptDampScf  (src/gscf/TruncatedOptDampScf.hh:239-409)
# --------------------------------------------------------------------------------------
def compute_damping_coeff(oda_s: float, oda_c: float, min_fock_prefactor: float = 0.05) -> float:
    """``TruncatedOptDampScf.hh:370-394`` verbatim control flow."""
    ignore_threshold = 1e-14
    if abs(oda_s) < ignore_threshold or abs(oda_c) < ignore_threshold:
        return 1.0
    if oda_s >= 0:
        return 1.0
    new = 1.0 if 2.0 * oda_c <= -oda_s else -oda_s / (2 * oda_c)
    assert 0.0 <= new <= 1.0
    if new < min_fock_prefactor:
        return min_fock_prefactor
    if new > 1 - min_fock_prefactor:
        return 1.0
    return new


def _trace_fprev_dcur(Fprev, orben, C, damping_coeff, n_alpha, n_beta):
    """``TruncatedOptDampScf.hh:291-347``."""
    nblk = Fprev.shape[0]
    restricted = nblk == 1
    if damping_coeff == 1:  # :309-319
        sum_a = float(np.sum(orben[0][:n_alpha]))
        if restricted:
            return 2.0 * sum_a
        return sum_a + float(np.sum(orben[1][:n_beta]))
    ca = C[0][:, :n_alpha]  # :328-336
    tr_aa = float(np.sum(ca * (Fprev[0] @ ca)))
    if restricted:
        return 2.0 * tr_aa
    cb = C[1][:, :n_beta]  # :342-344
    return tr_aa + float(np.sum(cb * (Fprev[1] @ cb)))


def truncated_opt_damp_scf(
    problem, F0, e1_0, e2_0, min_fock_prefactor=0.05, max_error_norm=5e-7, max_iter=100
) -> ScfResult:
    S = problem.S
    F, e1, e2 = np.array(F0, copy=True), e1_0, e2_0
    Fprev = None  # TruncatedOptDampScf.hh:61, :75
    e1_prev = e2_prev = None
    damp = 1.0  # :76
    n_iter, last_err = 0, float("nan")
    orben = C = None
    tE, tR, tC = [], [], []
    while not _check_convergence(n_iter, last_err, max_error_norm, max_iter):
        n_iter += 1
        # update_damping_coefficients :349-368
        if Fprev is None:
            damp = 1.0
        else:
            tr = _trace_fprev_dcur(Fprev, orben, C, damp, problem.n_alpha, problem.n_beta)
            oda_s = tr - e1_prev - 2.0 * e2_prev
            oda_c = e1 - tr + e2 + e2_prev
            damp = compute_damping_coeff(oda_s, oda_c, min_fock_prefactor)
            if damp == 1:
                Fprev = None
        # update_oda_diagmat :396-409
        if Fprev is None:
            Fdiag = F
        else:
            Fdiag = (1.0 - damp) * Fprev + damp * F
        orben, C = _eig_blocks(Fdiag, S)  # :253
        lam = damp  # weight of the current F in the matrix just diagonalised
        # keep/flip rule :273-277
        if damp > 0.5:
            Fprev, e1_prev, e2_prev = F, e1, e2
        else:
            damp = 1 - damp
        F, e1, e2 = problem.fock(C)  # :282
        last_err = _norm_frobenius(  # :285
            pulay_error(F, C, S, problem.n_alpha, problem.n_beta)
        )
        tE.append(e1 + e2)
        tR.append(last_err)
        tC.append(lam)
    return ScfResult(n_iter, True, last_err, orben, C, F, e1, e2, tE, tR, tC)
