"""CPU validation of gscf_b200/csrc/smallsolve.hpp (shared by the host code and the device kernels): the bordered DIIS
system with its LU solve against LAPACK dgesv on the oracle's matrix (PulayDiisScf.hh:326-393), ring-slot ordering, the
failure flag, and the TODA line search against the oracle's restatement (TruncatedOptDampScf.hh:370-394)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import compute_damping_coeff, diis_linear_system_matrix

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "cpp", "smallsolve_host.cc")
    out = os.path.join(HERE, "cpp", "libsmallsolve_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", src, "-o", out], check=True)
    L = C.CDLL(out)
    L.smallsolve_diis.restype = C.c_int
    L.smallsolve_diis.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.smallsolve_damping.restype = C.c_double
    L.smallsolve_damping.argtypes = [C.c_double, C.c_double, C.c_double]
    return L


def _solve(lib, gram, order):
    k = len(order)
    gram = np.ascontiguousarray(gram, dtype=np.float64)
    order = np.ascontiguousarray(order, dtype=np.int32)
    coeff = np.zeros(max(k, 1))
    ok = lib.smallsolve_diis(gram.ctypes.data, gram.shape[1], order.ctypes.data, k, coeff.ctypes.data)
    return ok, coeff[:k]


@pytest.mark.parametrize("k", [2, 3, 4, 8, 16])
def test_diis_solve_matches_dgesv_on_the_reference_matrix(lib, k):
    rng = np.random.default_rng(k)
    M = 16
    errs = rng.standard_normal((k, 40)) * np.logspace(0, -6, k)[:, None]   # errors shrink like a converging SCF
    # ring slots in arbitrary positions: order[i] = slot of the i-th oldest error
    order = rng.permutation(M)[:k]
    gram = np.full((M, M), np.nan)
    for i in range(k):
        for j in range(k):
            gram[order[i], order[j]] = errs[i] @ errs[j]
    ok, c = _solve(lib, gram, order)
    assert ok == 1
    # the reference's overlap vectors: entry i holds <e_i, e_i>, <e_i, e_{i-1}>, ... (newest first)
    overlaps = [[float(errs[i] @ errs[i - j]) for j in range(i + 1)] for i in range(k)]
    B = diis_linear_system_matrix(overlaps)
    rhs = np.zeros(k + 1)
    rhs[-1] = -1.0
    ref = sla.solve(B, rhs)[:k]
    np.testing.assert_allclose(c, ref, rtol=1e-9, atol=1e-12)
    assert abs(c.sum() - 1.0) < 1e-9            # the border row enforces sum c = 1


def test_diis_trivial_sizes_and_failure_flag(lib):
    gram = np.eye(16)
    ok, c = _solve(lib, gram, [3])
    assert ok == 1 and c.tolist() == [1.0]      # k = 1: c = [1] (PulayDiisScf.hh:366-369)
    ok, c = _solve(lib, gram, [])
    assert ok == 1 and len(c) == 0
    bad = np.zeros((16, 16))
    bad[0, 0] = bad[1, 1] = np.nan              # non-finite overlaps: ExcDiisStepFailed
    ok, _ = _solve(lib, bad, [0, 1])
    assert ok == 0


def test_damping_coefficient_matches_the_oracle(lib):
    rng = np.random.default_rng(3)
    cases = [(-1.0, 2.0), (1.0, 2.0), (-1.0, 0.4), (-1e-15, 1.0), (-1.0, 1e-15), (-0.02, 1.0), (-1.0, 0.5),
             (-1.9, 1.0), (-1.0, -3.0), (0.0, 0.0)]
    cases += [tuple(rng.standard_normal(2) * 10.0 ** rng.integers(-3, 3)) for _ in range(200)]
    for s, c in cases:
        for mp in (0.05, 0.2):
            assert lib.smallsolve_damping(s, c, mp) == compute_damping_coeff(s, c, mp), (s, c, mp)
