"""CPU validation of the numerical core of the divide & conquer eigensolver
(gscf_b200/csrc/dc_core.hpp: deflation + secular equation) through a serial host harness."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def dc():
    src = os.path.join(HERE, "cpp", "dc_host.cc")
    out = os.path.join(HERE, "cpp", "libdc_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", src, "-o", out], check=True)
    lib = C.CDLL(out)
    lib.dc_host_solve.restype = C.c_int
    lib.dc_host_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    def solve(d, e, leaf=8):
        n = len(d)
        d = np.ascontiguousarray(d, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        w = np.zeros(n)
        Z = np.zeros((n, n))
        st = np.zeros(2)
        lib.dc_host_solve(n, d.ctypes.data, e.ctypes.data if n > 1 else None, leaf, w.ctypes.data, Z.ctypes.data,
                          st.ctypes.data)
        return w, Z.T.copy(), st  # Z stored column-major

    return solve


def _check(d, e, w, Z, tol=60):
    n = len(d)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    wref = sla.eigh_tridiagonal(d, e, eigvals_only=True) if n > 1 else np.array(d)
    scale = max(np.abs(wref).max(), 1e-300)
    eps = 2.2e-16
    assert np.abs(w - wref).max() <= tol * eps * scale * max(1, np.sqrt(n))
    assert np.abs(Z.T @ Z - np.eye(n)).max() <= tol * eps * max(1, np.sqrt(n))
    assert np.abs(T @ Z - Z * w).max() <= tol * eps * scale * max(1, np.sqrt(n))


@pytest.mark.parametrize("n", [2, 3, 9, 17, 64, 100, 257])
def test_random_tridiagonal(dc, n):
    rng = np.random.default_rng(n)
    d = rng.standard_normal(n)
    e = rng.standard_normal(n - 1)
    w, Z, st = dc(d, e, leaf=4)
    _check(d, e, w, Z)
    assert st[1] < 40  # the rational iteration converges, bisection is only a safeguard


def test_structured_matrices(dc):
    # 1-2-1 Toeplitz (equally spaced-ish, heavy deflation at high levels)
    n = 200
    w, Z, st = dc(2 * np.ones(n), -np.ones(n - 1), leaf=8)
    _check(2 * np.ones(n), -np.ones(n - 1), w, Z)
    # Wilkinson W21+ (pairs of eigenvalues agreeing to ~1e-14) and glued Wilkinson (clusters)
    m = 10
    dW = np.abs(np.arange(-m, m + 1)).astype(float)
    eW = np.ones(2 * m)
    w, Z, st = dc(dW, eW, leaf=4)
    _check(dW, eW, w, Z)
    dG = np.tile(dW, 5)
    eG = np.concatenate([np.concatenate([eW, [1e-8]]) for _ in range(5)])[:-1]
    w, Z, st = dc(dG, eG, leaf=8)
    _check(dG, eG, w, Z)
    # graded matrix and a matrix with zero couplings (exact block splits -> K = 0 merges)
    n = 128
    dg = 10.0 ** (-np.arange(n) / 8.0)
    eg = 0.3 * np.sqrt(dg[:-1] * dg[1:])
    w, Z, st = dc(dg, eg, leaf=8)
    _check(dg, eg, w, Z)
    dz = np.arange(1.0, 65.0)
    ez = np.zeros(63)
    ez[::5] = 0.5
    w, Z, st = dc(dz, ez, leaf=8)
    _check(dz, ez, w, Z)
    # identical diagonal, tiny couplings (massive clustering)
    w, Z, st = dc(np.ones(96), 1e-12 * np.ones(95), leaf=8)
    _check(np.ones(96), 1e-12 * np.ones(95), w, Z)


def test_scf_like_spectrum(dc):
    """Tridiagonal form of a synthetic Fock-like matrix (what the SCF path feeds the solver)."""
    from oracle.synthetic import SyntheticDFProblem

    p = SyntheticDFProblem(192, 24, n_aux=8, seed=1)
    X = np.linalg.inv(np.linalg.cholesky(p.S))
    A = X @ p.H @ X.T
    H, Q = sla.hessenberg(A, calc_q=True)
    d, e = np.diag(H).copy(), np.diag(H, -1).copy()
    w, Z, st = dc(d, e, leaf=16)
    _check(d, e, w, Z)
    assert np.abs(w - np.linalg.eigvalsh(A)).max() < 1e-12


@pytest.mark.parametrize("leaf", [8, 16, 32])
def test_cascaded_noise_tridiagonal(dc, leaf):
    """What the tridiagonalisation of a numerically low-rank matrix (all ones) hands to the solver: a 2 x 2 block of
    norm n followed by blocks of noise whose magnitude drops by 1e-16 per block, down to 1e-176.  Deflation and the
    secular solver must keep the eigenvectors orthogonal (the GPU path additionally flushes entries below 1e-140 of the
    norm before its QL leaves: stedc.cu dc_scale_kernel)."""
    rng = np.random.default_rng(leaf)
    n = 96
    mag = np.array([10.0 ** (-16 * (i // 8)) for i in range(n)])
    d = rng.standard_normal(n) * mag
    e = rng.standard_normal(n - 1) * mag[1:]
    d[0], d[1], e[0] = 1.0, n - 1.0, np.sqrt(n - 1.0)
    w, Z, st = dc(d, e, leaf=leaf)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(Z.T @ Z - np.eye(n)).max() < 1e-13
    assert np.abs(T @ Z - Z * w).max() < 1e-12 * n
    wref = sla.eigh_tridiagonal(d, e, eigvals_only=True)
    assert np.abs(np.sort(w) - wref).max() < 1e-12 * n
