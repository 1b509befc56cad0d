"""Synthetic density-fitted closed-/open-shell problem (TEST ORACLE side).

gscf contains no integrals (BASELINE.json north_star); the benchmark problem is
the deterministic generator specified in SURVEY.md section 8(d):

    u(seed, k):   z = seed + (k+1) * 0x9E3779B97F4A7C15            (mod 2^64)
                  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9
                  z = (z ^ (z >> 27)) * 0x94D049BB133111EB
                  z ^= z >> 31                                      (splitmix64 finaliser)
                  u = (z >> 11) * 2^-52 - 1            in [-1, 1)
    symrand(seed)_ij = u(seed, hi*(hi+1)/2 + lo),  hi = max(i,j), lo = min(i,j)
    S   = I + (0.25/sqrt(n)) * symrand(1000 p + 1), diag forced to 1
    H   = diag(h) + (0.5/sqrt(n)) * symrand(1000 p + 2),  h_i = 2 i/n - 1 + 0.5 [i >= n_occ]
    B_Q = (g/sqrt(n n_aux)) * symrand(1000 p + 10 + Q),  Q < n_aux,  g = 1
    restricted:   D = C_o C_o^T, gamma_Q = 2 tr(B_Q D), J = sum_Q gamma_Q B_Q,
                  K = sum_Q B_Q D B_Q, F = H + J - K, E1e = 2 tr(H D), E2e = tr(J D) - tr(K D)
    unrestricted: gamma_Q = tr(B_Q (Da + Db)), F^s = H + J - K(D^s),
                  E1e = tr(H (Da+Db)), E2e = 1/2 [tr(J (Da+Db)) - tr(K^a Da) - tr(K^b Db)]
    guess        = core guess: eigenvectors of H against S.

The same generator is implemented on the device in ``gscf_b200/csrc/fock_df.cu``;
``tests/test_synthetic.py`` checks the two bit for bit.
"""
from __future__ import annotations

import numpy as np

from .scf import generalised_eigh

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix_u(seed: int, k: np.ndarray) -> np.ndarray:
    """u(seed, k) in [-1, 1), vectorised over k (uint64 array)."""
    with np.errstate(over="ignore"):
        k = k.astype(np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + (k + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * 2.0**-52 - 1.0


def symrand(seed: int, n: int) -> np.ndarray:
    i, j = np.meshgrid(np.arange(n, dtype=np.uint64), np.arange(n, dtype=np.uint64), indexing="ij")
    hi = np.maximum(i, j)
    lo = np.minimum(i, j)
    return splitmix_u(seed, hi * (hi + np.uint64(1)) // np.uint64(2) + lo)


def _k_from_y(Y: np.ndarray) -> np.ndarray:
    """K = sum_Q Y_Q Y_Q^T as one dgemm (Y: (aux, n, o))."""
    q, n, o = Y.shape
    Yt = np.ascontiguousarray(Y.transpose(1, 0, 2)).reshape(n, q * o)
    return Yt @ Yt.T


class SyntheticDFProblem:
    """Fock-like plugin (``FocklikeMatrix_i``) for the synthetic DF problem."""

    def __init__(self, n_bas: int, n_alpha: int, n_beta: int | None = None, n_aux: int | None = None,
                 seed: int = 0, g: float = 1.0, unrestricted: bool | None = None, summation: str = "pairwise"):
        n = n_bas
        # how the energy traces are summed: numpy's pairwise reduction (default) or a plain left-to-right sum.  Both are
        # legitimate fp64 evaluations of the same formula; the second exists to show how far the damped TODA branch moves
        # under a mere change of summation order (tests/test_toda_sensitivity_cpu.py)
        assert summation in ("pairwise", "sequential")
        self._sum = (lambda a: float(np.sum(a))) if summation == "pairwise" else (lambda a: float(np.cumsum(a.ravel())[-1]))
        self.n_bas = n
        self.n_alpha = n_alpha
        self.n_beta = n_alpha if n_beta is None else n_beta
        self.unrestricted = (self.n_alpha != self.n_beta) if unrestricted is None else unrestricted
        self.n_aux = n_aux if n_aux is not None else max(1, n // 16)
        self.seed = seed
        p = seed
        S = np.eye(n) + (0.25 / np.sqrt(n)) * symrand(1000 * p + 1, n)
        np.fill_diagonal(S, 1.0)
        h = 2.0 * np.arange(n) / n - 1.0 + 0.5 * (np.arange(n) >= self.n_alpha)
        H = (0.5 / np.sqrt(n)) * symrand(1000 * p + 2, n)
        H[np.diag_indices(n)] += h
        self.S, self.H = S, H
        scale = g / np.sqrt(n * self.n_aux)
        self.B = np.stack([scale * symrand(1000 * p + 10 + Q, n) for Q in range(self.n_aux)])

    # -- plugin interface -------------------------------------------------------------
    def core_guess(self) -> np.ndarray:
        w, v = generalised_eigh(self.H, self.S)
        return v

    def fock(self, C: np.ndarray):
        C = np.asarray(C)
        if C.ndim == 2:
            C = C[None]
        B, H = self.B, self.H
        if not self.unrestricted:
            co = C[0][:, : self.n_alpha]
            D = co @ co.T
            Y = B @ co  # (aux, n, o)
            gamma = 2.0 * np.einsum("qia,ia->q", Y, co)
            J = np.tensordot(gamma, B, 1)
            K = _k_from_y(Y)
            F = H + J - K
            e1 = 2.0 * self._sum(H * D)
            e2 = self._sum(J * D) - self._sum(K * D)
            return F[None], e1, e2
        ca = C[0][:, : self.n_alpha]
        cb = C[1 if C.shape[0] > 1 else 0][:, : self.n_beta]
        Da, Db = ca @ ca.T, cb @ cb.T
        Ya, Yb = B @ ca, B @ cb
        gamma = np.einsum("qia,ia->q", Ya, ca) + np.einsum("qia,ia->q", Yb, cb)
        J = np.tensordot(gamma, B, 1)
        Ka = _k_from_y(Ya)
        Kb = _k_from_y(Yb)
        Dt = Da + Db
        e1 = self._sum(H * Dt)
        e2 = 0.5 * (self._sum(J * Dt) - self._sum(Ka * Da) - self._sum(Kb * Db))
        return np.stack([H + J - Ka, H + J - Kb]), e1, e2
