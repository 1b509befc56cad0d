"""Kernel-level parity through the C ABI against numpy on the same seeded inputs (GPU)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch

    from gscf_b200 import lib as L

    assert torch.cuda.is_available()
    return torch, L


def dev(torch, a):
    """numpy array -> device tensor holding the same bytes (flat, column-major preserved)."""
    return torch.from_numpy(np.ascontiguousarray(a).ravel().copy()).cuda()


def colmajor(a, ld=None):
    """2-D numpy -> flat column-major buffer with leading dimension ld (zero padded)."""
    m, n = a.shape
    ld = m if ld is None else ld
    buf = np.zeros((n, ld))
    buf[:, :m] = a.T
    return buf.ravel()


def from_colmajor(buf, m, n, ld):
    return np.asarray(buf).reshape(n, ld)[:, :m].T.copy()


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("m,n,k", [(14, 14, 2), (64, 64, 64), (129, 67, 33), (257, 300, 128), (1024, 1024, 96),
                                   (1, 1, 1), (5, 3, 0)])
def test_dgemm(env, ta, tb, m, n, k):
    torch, L = env
    rng = np.random.default_rng(m * 1000 + n * 10 + k + ta * 2 + tb)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    Cm = rng.standard_normal((m, n))
    lda = A.shape[0] + (A.shape[0] % 2) + 2
    ldb = B.shape[0] + 3  # odd leading dimension: exercises the 8-byte cp.async path
    ldc = m + 1
    dA, dB, dC = dev(torch, colmajor(A, lda)), dev(torch, colmajor(B, ldb)), dev(torch, colmajor(Cm, ldc))
    alpha, beta = 1.5, -0.25
    L.check(L.lib.gscf_b200_dgemm(ta, tb, m, n, k, alpha, dA.data_ptr(), lda, 0, dB.data_ptr(), ldb, 0, beta,
                                  dC.data_ptr(), ldc, 0, 1, None))
    torch.cuda.synchronize()
    got = from_colmajor(dC.cpu().numpy(), m, n, ldc)
    opA = A.T if ta else A
    opB = B.T if tb else B
    ref = alpha * (opA @ opB) + beta * Cm
    tol = 1e-13 * max(1, k) * (1 + np.abs(ref).max())
    assert np.abs(got - ref).max() < tol


def test_dgemm_batched_and_large_tile(env):
    torch, L = env
    rng = np.random.default_rng(7)
    batch, m, n, k = 3, 1536, 1408, 160  # large-tile configuration (>= 120 128x128 tiles)
    A = rng.standard_normal((batch, k, m))  # column-major m x k each
    B = rng.standard_normal((batch, n, k))  # column-major k x n each
    dA, dB = dev(torch, A), dev(torch, B)
    dC = torch.zeros(batch * m * n, dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_dgemm(0, 0, m, n, k, 1.0, dA.data_ptr(), m, m * k, dB.data_ptr(), k, k * n, 0.0,
                                  dC.data_ptr(), m, m * n, batch, None))
    torch.cuda.synchronize()
    got = dC.cpu().numpy().reshape(batch, n, m)
    for b in range(batch):
        ref = (A[b].T @ B[b].T).T  # stored column-major == transpose
        assert np.abs(got[b] - ref).max() < 1e-11


@pytest.mark.parametrize("ta,tb,m,n,k,batch", [
    (0, 0, 1536, 1400, 160, 2),    # NN: A through the 4-D row-group map, B through its k-major map, ragged n
    (1, 0, 1300, 1290, 130, 1),    # TN: both k-major, ragged m / n / k (zero fill by the TMA unit)
    (1, 1, 1400, 1403, 64, 2),     # TT, strided batch (last map coordinate = batch entry), n not a multiple of 8
    (0, 1, 1541, 1408, 100, 2),    # NT: both operands through 4-D row-group maps, ragged m, k not a multiple of the k-tile
    (1, 0, 1024, 1024, 100, 2),    # TN, k not a multiple of the k-tile
    (0, 1, 2048, 2048, 512, 1),    # NT, all tiles interior
])
def test_dgemm_tensor_map_path(env, ta, tb, m, n, k, batch):
    """Operands staged by tensor-map TMA (cp.async.bulk.tensor, `dgemm_dmma_tm_kernel`; the 128 x 128 tile configuration):
    result against numpy, NaN in every padding row (rows the maps may touch must not reach stored results), and proof
    that the path ran (the launch counter of that kernel moves)."""
    import os

    torch, L = env
    if os.environ.get("GSCF_B200_GEMM_TMA", "3") not in ("3", "4"):
        pytest.skip("tensor-map path switched off by GSCF_B200_GEMM_TMA")
    rng = np.random.default_rng(m + 3 * n + 7 * k + ta + 2 * tb)
    ra, ca = (k, m) if ta else (m, k)
    rb, cb = (n, k) if tb else (k, n)
    lda, ldb, ldc = (ra + 7) // 8 * 8 + 2, (rb + 7) // 8 * 8, m + 1   # even, >= rows rounded up to 8
    sA, sB, sC = lda * ca + 4, ldb * cb, ldc * n
    A = rng.standard_normal((batch, ra, ca))
    B = rng.standard_normal((batch, rb, cb))
    Cm = rng.standard_normal((batch, m, n))
    bufA, bufB, bufC = np.full(batch * sA, np.nan), np.full(batch * sB, np.nan), np.zeros(batch * sC)
    for b in range(batch):
        va = bufA[b * sA: b * sA + lda * ca].reshape(ca, lda)
        va[:, :ra] = A[b].T
        vb = bufB[b * sB: b * sB + ldb * cb].reshape(cb, ldb)
        vb[:, :rb] = B[b].T
        bufC[b * sC: b * sC + ldc * n] = colmajor(Cm[b], ldc)
    dA, dB, dC = dev(torch, bufA), dev(torch, bufB), dev(torch, bufC)
    alpha, beta = -0.75, 0.5
    before = L.lib.gscf_b200_tensor_map_launch_count()
    L.check(L.lib.gscf_b200_dgemm(ta, tb, m, n, k, alpha, dA.data_ptr(), lda, sA, dB.data_ptr(), ldb, sB, beta,
                                  dC.data_ptr(), ldc, sC, batch, None))
    torch.cuda.synchronize()
    assert L.lib.gscf_b200_tensor_map_launch_count() == before + 1, "the tensor-map kernel did not run"
    got = dC.cpu().numpy()
    for b in range(batch):
        opA = A[b].T if ta else A[b]
        opB = B[b].T if tb else B[b]
        ref = alpha * (opA @ opB) + beta * Cm[b]
        gb = from_colmajor(got[b * sC: b * sC + ldc * n], m, n, ldc)
        assert np.abs(gb - ref).max() < 1e-13 * max(1, k) * (1 + np.abs(ref).max())


def test_dgemm_randomised_sweep_tensor_maps_on_small_tiles():
    """The randomised GEMM sweep in a child process with GSCF_B200_GEMM_TMA=4: the tensor-map kernels also on the
    64 x 64 tiles, where the default leaves them off (measured slower) - most shapes of the sweep use those tiles."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env2 = dict(os.environ, GSCF_B200_GEMM_TMA="4", GSCF_B200_TMAP_STRICT="0")  # also products smaller than one box
    r = subprocess.run([sys.executable, "scripts/gemm_fuzz.py", "150", "900"], cwd=root, env=env2, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "failures 0" in r.stdout


@pytest.mark.parametrize("n", [1, 2, 3, 14, 31, 32, 33, 64, 100, 128, 160])
def test_syev_jacobi(env, n):
    torch, L = env
    rng = np.random.default_rng(n)
    batch = 3
    ld = n + (n % 2)
    mats = []
    for b in range(batch):
        a = rng.standard_normal((n, n))
        a = a + a.T
        if b == 1 and n > 4:  # degenerate / clustered spectrum
            q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            ev = np.repeat(np.arange(1, n // 2 + 2), 2)[:n].astype(float)
            a = (q * ev) @ q.T
            a = 0.5 * (a + a.T)
        mats.append(a)
    buf = np.concatenate([colmajor(np.tril(a) + 7.0 * np.triu(np.ones_like(a), 1), ld) for a in mats])
    dA = dev(torch, buf)
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    L.check(L.lib.gscf_b200_syev(L.EIG_JACOBI, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld,
                                 ld * n, batch, None, 0, info.data_ptr(), None))
    torch.cuda.synchronize()
    assert info.cpu().numpy().tolist() == [0] * batch
    w = dw.cpu().numpy().reshape(batch, n)
    Z = dZ.cpu().numpy().reshape(batch, n, ld)
    for b, a in enumerate(mats):
        wref = np.linalg.eigvalsh(a)
        scale = max(1.0, np.abs(wref).max())
        assert np.abs(w[b] - wref).max() < 1e-12 * scale * max(1, n / 8)
        z = Z[b][:, :n].T
        assert np.abs(z.T @ z - np.eye(n)).max() < 1e-12
        assert np.abs(a @ z - z * w[b]).max() < 2e-12 * scale * max(1, n / 8)


def test_frob_multidot_and_axpby(env):
    torch, L = env
    rng = np.random.default_rng(3)
    for (n, cols, ld, count, batch) in [(14, 14, 16, 3, 1), (127, 254, 128, 5, 2), (256, 256, 256, 9, 1),
                                        (1024, 1024, 1024, 8, 1)]:
        nslots = count + 1
        ring = rng.standard_normal((batch, nslots, cols, ld))
        enew = rng.standard_normal((batch, cols, ld))
        slots = np.array(rng.permutation(nslots)[:count], dtype=np.int32)
        dR, dE = dev(torch, ring), dev(torch, enew)
        out = torch.zeros(batch * 16, dtype=torch.float64, device="cuda")
        wsz = L.lib.gscf_b200_frob_multidot_worksize(count, n, cols, batch)
        work = torch.zeros(max(1, wsz // 8), dtype=torch.float64, device="cuda")
        L.check(L.lib.gscf_b200_frob_multidot(dE.data_ptr(), dR.data_ptr(), cols * ld,
                                              slots.ctypes.data_as(C.POINTER(C.c_int)), count, n, cols, ld,
                                              cols * ld, nslots * cols * ld, batch, out.data_ptr(), 16,
                                              work.data_ptr(), wsz, None))
        coeff = rng.standard_normal((batch, 16))
        dcoef = dev(torch, coeff)
        res = torch.zeros(batch * cols * ld, dtype=torch.float64, device="cuda")
        L.check(L.lib.gscf_b200_axpby_n(res.data_ptr(), cols * ld, dR.data_ptr(), cols * ld, nslots * cols * ld,
                                        slots.ctypes.data_as(C.POINTER(C.c_int)), count, dcoef.data_ptr(), 16, n, cols,
                                        ld, batch, None))
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(batch, 16)
        gres = res.cpu().numpy().reshape(batch, cols, ld)
        for b in range(batch):
            for j, s in enumerate(slots):
                ref = np.sum(enew[b][:, :n] * ring[b, s][:, :n])
                assert abs(got[b, j] - ref) < 1e-12 * n * cols
            ref = sum(coeff[b, j] * ring[b, s][:, :n] for j, s in enumerate(slots))
            assert np.abs(gres[b][:, :n] - ref).max() < 1e-13 * count * 10


def test_diis_solve_matches_dgesv(env):
    torch, L = env
    from oracle import diis_linear_system_matrix

    rng = np.random.default_rng(11)
    M = 8
    batch = 5
    for k in [2, 3, 5, 8]:
        errs = rng.standard_normal((batch, M, 40)) * np.logspace(0, -6, M)[None, :, None]
        gram = np.einsum("bik,bjk->bij", errs, errs)
        order = np.array(rng.permutation(M)[:k], dtype=np.int32)
        dG = dev(torch, gram)
        coef = torch.zeros(batch * 16, dtype=torch.float64, device="cuda")
        fail = torch.zeros(batch, dtype=torch.int32, device="cuda")
        L.check(L.lib.gscf_b200_diis_solve(dG.data_ptr(), M, M * M, order.ctypes.data_as(C.POINTER(C.c_int)), k,
                                           coef.data_ptr(), 16, fail.data_ptr(), batch, None))
        torch.cuda.synchronize()
        c = coef.cpu().numpy().reshape(batch, 16)
        assert fail.cpu().numpy().sum() == 0
        for b in range(batch):
            ov = [[gram[b, order[i], order[i - j]] for j in range(i + 1)] for i in range(k)]
            B = diis_linear_system_matrix(ov)
            rhs = np.zeros(k + 1)
            rhs[k] = -1
            ref = np.linalg.solve(B, rhs)[:k]
            assert np.abs(c[b, :k] - ref).max() < 1e-9 * max(1, np.abs(ref).max())
            assert abs(c[b, :k].sum() - 1) < 1e-10
    # singular system -> failure flag (ExcDiisStepFailed)
    gram = np.zeros((1, M, M))
    dG = dev(torch, gram)
    order = np.arange(3, dtype=np.int32)
    coef = torch.zeros(16, dtype=torch.float64, device="cuda")
    fail = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check(L.lib.gscf_b200_diis_solve(dG.data_ptr(), M, M * M, order.ctypes.data_as(C.POINTER(C.c_int)), 3,
                                       coef.data_ptr(), 16, fail.data_ptr(), 1, None))
    torch.cuda.synchronize()
    assert fail.cpu().numpy()[0] == 1


@pytest.mark.parametrize("n,o", [(14, 2), (128, 16), (200, 25)])
def test_density_and_pulay_error(env, n, o):
    torch, L = env
    from oracle import pulay_error

    rng = np.random.default_rng(n)
    ld = L.lib.gscf_b200_padded_ld(n)
    S = rng.standard_normal((n, n)); S = S @ S.T / n + np.eye(n)
    F = rng.standard_normal((n, n)); F = F + F.T
    Cm = rng.standard_normal((n, n))
    dS, dF, dC = dev(torch, colmajor(S, ld)), dev(torch, colmajor(F, ld)), dev(torch, colmajor(Cm, ld))
    dD = torch.zeros(ld * n, dtype=torch.float64, device="cuda")
    dE = torch.zeros(ld * n, dtype=torch.float64, device="cuda")
    wsz = L.lib.gscf_b200_pulay_error_worksize(n, o, 1)
    work = torch.zeros(wsz // 8, dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_density(n, o, dC.data_ptr(), ld, 0, dD.data_ptr(), ld, 0, 1, None))
    L.check(L.lib.gscf_b200_pulay_error(n, o, 2.0, dS.data_ptr(), ld, 0, dF.data_ptr(), ld, 0, dC.data_ptr(), ld, 0,
                                        dE.data_ptr(), ld, 0, work.data_ptr(), 0, 1, None))
    torch.cuda.synchronize()
    D = from_colmajor(dD.cpu().numpy(), n, n, ld)
    E = from_colmajor(dE.cpu().numpy(), n, n, ld)
    assert np.abs(D - Cm[:, :o] @ Cm[:, :o].T).max() < 1e-12 * o
    ref = pulay_error(F[None], Cm[None], S, o, o)[0]
    assert np.abs(E - ref).max() < 1e-11 * np.abs(ref).max()
    # the fused product stores every strictly lower entry twice with opposite signs: exactly antisymmetric, zero diagonal
    assert np.array_equal(E, -E.T) and not np.any(np.diag(E))


@pytest.mark.parametrize("n,o,batch", [(70, 9, 5), (257, 33, 3)])
def test_pulay_error_batched(env, n, o, batch):
    """The strided-batch form of gscf_b200_pulay_error (one S shared, F / C / E / work per entry), edge tiles in both
    tile sizes, against pulay_error.cc:25-44 restated in the oracle."""
    torch, L = env
    from oracle import pulay_error

    rng = np.random.default_rng(7 * n + batch)
    ld = L.lib.gscf_b200_padded_ld(n)
    S = rng.standard_normal((n, n)); S = S @ S.T / n + np.eye(n)
    Fs = [(lambda a: a + a.T)(rng.standard_normal((n, n))) for _ in range(batch)]
    Cs = [rng.standard_normal((n, n)) for _ in range(batch)]
    dS = dev(torch, colmajor(S, ld))
    dF = dev(torch, np.concatenate([colmajor(f, ld) for f in Fs]))
    dC = dev(torch, np.concatenate([colmajor(c, ld) for c in Cs]))
    dE = torch.full((batch * ld * n,), 7.0, dtype=torch.float64, device="cuda")
    wsz = L.lib.gscf_b200_pulay_error_worksize(n, o, batch)
    work = torch.zeros(wsz // 8, dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_pulay_error(n, o, 1.0, dS.data_ptr(), ld, 0, dF.data_ptr(), ld, ld * n, dC.data_ptr(), ld,
                                        ld * n, dE.data_ptr(), ld, ld * n, work.data_ptr(), wsz // 8 // batch, batch, None))
    torch.cuda.synchronize()
    Eall = dE.cpu().numpy()
    for b in range(batch):
        E = from_colmajor(Eall[b * ld * n:(b + 1) * ld * n], n, n, ld)
        ref = 0.5 * pulay_error(Fs[b][None], Cs[b][None], S, o, o)[0]  # the oracle's closed-shell factor 2 taken out
        assert np.abs(E - ref).max() < 1e-11 * np.abs(ref).max()
        assert np.array_equal(E, -E.T)


def test_measure_peaks_runs(env):
    torch, L = env
    a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
    L.check(L.lib.gscf_b200_measure_peaks(C.byref(a), C.byref(b), C.byref(c)))
    assert a.value > 1.0 and b.value > 1.0 and c.value > 500.0


def _syev_check(torch, L, method, mats, n, tolscale=1.0):
    batch = len(mats)
    ld = L.lib.gscf_b200_padded_ld(n)
    buf = np.concatenate([colmajor(np.tril(a) + 3.0 * np.triu(np.ones_like(a), 1), ld) for a in mats])
    dA = dev(torch, buf)
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.full((batch,), 7, dtype=torch.int32, device="cuda")
    wsz = L.lib.gscf_b200_syev_worksize(n, batch, method)
    work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_syev(method, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld, ld * n,
                                 batch, work.data_ptr(), wsz, info.data_ptr(), None))
    torch.cuda.synchronize()
    assert info.cpu().numpy().tolist() == [0] * batch
    w = dw.cpu().numpy().reshape(batch, n)
    Z = dZ.cpu().numpy().reshape(batch, n, ld)
    for b, a in enumerate(mats):
        wref = np.linalg.eigvalsh(a)
        scale = max(1.0, np.abs(wref).max())
        z = Z[b][:, :n].T
        assert np.abs(w[b] - wref).max() < tolscale * 1e-13 * scale * n
        assert np.abs(z.T @ z - np.eye(n)).max() < tolscale * 1e-13 * n
        assert np.abs(a @ z - z * w[b]).max() < tolscale * 1e-13 * scale * n


@pytest.mark.parametrize("n", [2, 5, 33, 64, 65, 130, 200, 257, 512, 1024])
def test_syev_tridiag_dc(env, n):
    torch, L = env
    rng = np.random.default_rng(n + 5)
    a = rng.standard_normal((n, n))
    a = a + a.T
    mats = [a]
    if n <= 257:
        # clustered spectrum (forces deflation through rotations) and a diagonal-ish matrix
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        ev = np.repeat(np.arange(1, n // 3 + 2), 3)[:n].astype(float)
        b = (q * ev) @ q.T
        mats.append(0.5 * (b + b.T))
        mats.append(np.diag(np.arange(n, dtype=float)) + 1e-9 * a)
    _syev_check(torch, L, L.EIG_TRIDIAG_DC, mats, n)


@pytest.mark.parametrize("n,k,batch", [(200, 25, 3), (512, 64, 2), (1024, 128, 1), (130, 129, 2), (300, 1, 2)])
def test_syevx_partial_spectrum(env, n, k, batch):
    """gscf_b200_syevx (n_eigenpairs = k < n, ScfBase.hh:87): all eigenvalues, the k lowest eigenvectors - through the
    partial top-level merge of the divide & conquer.  Random, clustered (heavy deflation: the rank of a root has to count
    the deflated eigenvalues below it) and nearly diagonal matrices."""
    torch, L = env
    rng = np.random.default_rng(n + k)
    mats = []
    for b in range(batch):
        a = rng.standard_normal((n, n))
        a = a + a.T
        if b == 1:
            qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
            ev = np.repeat(np.arange(1, n // 3 + 2), 3)[:n].astype(float)
            a = (qm * ev) @ qm.T
            a = 0.5 * (a + a.T)
        if b == 2:
            a = np.diag(np.arange(n, dtype=float)[::-1].copy()) + 1e-9 * a
        mats.append(a)
    ld = L.lib.gscf_b200_padded_ld(n)
    dA = dev(torch, np.concatenate([colmajor(a, ld) for a in mats]))
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.full((batch,), 7, dtype=torch.int32, device="cuda")
    wsz = L.lib.gscf_b200_syev_worksize(n, batch, L.EIG_TRIDIAG_DC)
    work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_syevx(L.EIG_TRIDIAG_DC, n, k, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld,
                                  ld * n, batch, work.data_ptr(), wsz, info.data_ptr(), None))
    torch.cuda.synchronize()
    assert info.cpu().numpy().tolist() == [0] * batch
    w = dw.cpu().numpy().reshape(batch, n)
    Z = dZ.cpu().numpy().reshape(batch, n, ld)
    for b, a in enumerate(mats):
        wref = np.linalg.eigvalsh(a)
        scale = max(1.0, np.abs(wref).max())
        z = Z[b][:k, :n].T  # n x k
        assert np.abs(w[b] - wref).max() < 1e-13 * scale * n
        assert np.abs(z.T @ z - np.eye(k)).max() < 1e-13 * n
        assert np.abs(a @ z - z * w[b][:k]).max() < 1e-13 * scale * n


def test_syev_tridiag_dc_fock_like(env):
    """The matrices the SCF feeds the solver: X^T F X of the synthetic problem, n = 384."""
    torch, L = env
    from oracle.synthetic import SyntheticDFProblem

    p = SyntheticDFProblem(384, 48, n_aux=4, seed=2)
    w, v = np.linalg.eigh(p.S)
    X = (v / np.sqrt(w)) @ v.T
    F, _, _ = p.fock(p.core_guess()[None])
    A = X @ F[0] @ X
    A = 0.5 * (A + A.T)
    _syev_check(torch, L, L.EIG_TRIDIAG_DC, [A, X @ p.H @ X], 384)


@pytest.mark.parametrize("n", [17, 64, 129, 256, 700])
def test_syev_special_matrices(env, n):
    """Zero / identity / rank-one / extreme-norm / glued-Wilkinson / graded matrices in one batch (scripts/eig_stress.py
    is the long version).  The all-ones matrix is the hard one: its tridiagonal form carries a cascade of noise down to
    1e-160 whose squares are subnormal (reflector norms, QL rotations); 1e+-150 matrices need dsyev's scaling, and the
    deflation tolerances of the D&C merges need dstedc's scaling of T to unit norm."""
    torch, L = env
    from scripts.eig_stress import kinds, run  # repo root is on sys.path (tests/conftest.py)

    names, mats = zip(*kinds(n, np.random.default_rng(n)))
    info, W, Z = run(L.EIG_AUTO, n, list(mats))
    for b, (name, a) in enumerate(zip(names, mats)):
        wref = np.linalg.eigvalsh(a)
        scale = max(np.abs(wref).max(), np.finfo(float).tiny)
        z = Z[b].T
        tol = 2e-13 * n
        assert info[b] == 0, name
        assert np.abs(W[b] - wref).max() / scale < tol, name
        assert np.abs(z.T @ z - np.eye(n)).max() < tol, name
        assert np.abs(a @ z - z * W[b]).max() / scale < tol, name


@pytest.mark.parametrize("n,batch,envs", [
    (333, 2, {"GSCF_B200_TF_GLOBAL": "1"}),                                        # trailing matrix in global memory
    (600, 1, {"GSCF_B200_TF_GLOBAL": "1", "GSCF_B200_TF_UNBLOCKED_BELOW": "64"}),   # blocked (deferred updates)
    (257, 3, {"GSCF_B200_TF_GLOBAL": "1", "GSCF_B200_TF_UNBLOCKED_BELOW": "100"}),  # blocked, odd order, switch-over
    (128, 40, {"GSCF_B200_TF_SOLO": "0"}),                                          # one cluster per matrix
    (512, 12, {"GSCF_B200_TF_CLUSTER_BARRIER": "0"}),                               # clusters with the L2 counter exchange
    (130, 5, {}),                                                                   # one CTA per matrix (SOLO, shared memory)
    (101, 6, {}),                                                                   # one CTA per matrix, registers
    (128, 7, {}),                                                                   # registers: 128-row layout, trailing 64 x 64 block handed to the 64-row layout
    (128, 7, {"GSCF_B200_TF_SOLO_SPLIT": "0"}),                                     # registers: 128-row layout to the end
    (100, 9, {}),                                                                   # hand-over at column 36
    (64, 11, {}),                                                                   # 64-row register layout directly
    (33, 5, {}),                                                                    # 64-row register layout, odd order
    (128, 40, {"GSCF_B200_DC_SECULAR": "thread"}),                                  # secular equation: one thread per root
    (512, 12, {"GSCF_B200_DC_SECULAR": "thread"}),                                  # ... up to 256 poles, a warp per root above
    (257, 9, {"GSCF_B200_DC_SECULAR": "thread", "GSCF_B200_DC_LEAF": "32"}),
    (128, 40, {"GSCF_B200_DC_LEAF": "32"}),                                         # D&C leaves of 32 (the default above 1024 matrices)
    (128, 4, {"GSCF_B200_TF_SOLO_REG": "0"}),                                       # n <= 128 through the shared-memory SOLO
    (512, 16, {}),                                                                  # clusters of 4 in L2, lower triangle only (SYMH)
    (512, 16, {"GSCF_B200_TF_CLUSTER_BARRIER": "0"}),                               # SYMH with the L2 counter exchange
    (512, 16, {"GSCF_B200_TF_SYMH_OCC": "1"}),                                      # SYMH, one CTA per SM build
    (512, 20, {"GSCF_B200_TF_BATCH_GLOBAL": "2"}),                                  # SYMH, clusters of 2
    (401, 30, {}),                                                                  # SYMH, odd order
    (1000, 20, {}),                                                                 # SYMH, four rows per thread
    (512, 16, {"GSCF_B200_TF_SYMH": "0"}),                                          # clusters of 4 in L2, full-column sweep
    (1024, 2, {"GSCF_B200_TRD_FUSED": "0"}),                                        # three-kernels-per-column fall-back
    (1024, 1, {"GSCF_B200_BT_PAIR": "0"}),                                          # back-transformation block by block (no merged pairs)
    (700, 3, {"GSCF_B200_BT_PAIR": "0"}),
    (385, 2, {}),                                                                   # three blocks: one merged pair + a single block
    (300, 20, {"GSCF_B200_BT_PAIR": "2"}),                                          # merged pairs forced for a batch (64-wide blocks)
    (128, 40, {"GSCF_B200_BT_SMALL": "2"}),                                         # fused back-transformation, lanes along the rows
    (128, 12, {"GSCF_B200_BT_SMALL": "0"}),                                         # blocked back-transformation for small matrices
    (100, 7, {"GSCF_B200_BT_SMALL": "0"}),
    (1024, 1, {"GSCF_B200_TF_SOLO_TAIL": "0"}),                                     # all-SM register kernel to the last column
    (1024, 2, {}),                                                                  # two matrices: tail of each in one CTA
    (2048, 1, {}),                                                                  # L2 kernel, hand-over to shared memory, then to one CTA
    (1801, 1, {}),                                                                  # shortest L2 phase (136 columns), odd order
    (2601, 1, {}),                                                                  # deferred updates, unblocked, hand-over
    (700, 1, {"GSCF_B200_TF_HANDOVER": "256"}),                                     # hand-over out of the register kernel
    (700, 1, {"GSCF_B200_TF_GLOBAL": "1", "GSCF_B200_TF_UNBLOCKED_BELOW": "300",
              "GSCF_B200_TF_HANDOVER": "200"}),                                     # hand-over out of the blocked kernel
    (1024, 1, {"GSCF_B200_TF_PHASES": "1"}),                                        # all SMs -> 16-CTA cluster -> one CTA
    (700, 2, {"GSCF_B200_TF_PHASES": "1", "GSCF_B200_TF_PHASE_CLUSTER": "0"}),      # hand-over to the single-CTA tail
])
def test_syev_fused_variants(env, n, batch, envs):
    """Every storage / exchange variant of the fused tridiagonalisation (trd_fused.cuh, trd_fused_blk.cuh)."""
    import os
    import subprocess
    import sys

    child_env = dict(os.environ)
    child_env.update(envs)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "syev_variant_main.py"), str(n), str(batch)],
                       cwd=root, env=child_env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("n,method", [(40, 1), (100, 2), (300, 2), (1024, 2)])
def test_syev_nan_input_is_contained(env, n, method):
    """Garbage in one matrix of a batch (a NaN entry) must neither hang a spinning kernel nor send an index-building
    kernel out of bounds (the D&C rank sorts use a total order on the bit patterns); the matrix is flagged through
    info or non-finite eigenvalues and its neighbours in the batch are still solved correctly."""
    torch, L = env
    rng = np.random.default_rng(n)
    batch = 1 if n > 512 else 3
    ld = L.lib.gscf_b200_padded_ld(n)
    a = rng.standard_normal((n, n))
    a = a + a.T
    buf = np.zeros((batch, n, ld))
    buf[:, :, :n] = a
    buf[0, 5, 7] = buf[0, 7, 5] = np.nan
    dA = dev(torch, buf.ravel())
    dw = torch.zeros(batch * n, dtype=torch.float64, device="cuda")
    dZ = torch.zeros(batch * ld * n, dtype=torch.float64, device="cuda")
    info = torch.zeros(batch, dtype=torch.int32, device="cuda")
    wsz = L.lib.gscf_b200_syev_worksize(n, batch, method)
    work = torch.zeros(max(1, wsz // 8 + 1), dtype=torch.float64, device="cuda")
    L.check(L.lib.gscf_b200_syev(method, n, dA.data_ptr(), ld, ld * n, dw.data_ptr(), n, dZ.data_ptr(), ld, ld * n,
                                 batch, work.data_ptr(), wsz, info.data_ptr(), None))
    torch.cuda.synchronize()  # an illegal address would surface here
    w = dw.cpu().numpy().reshape(batch, n)
    flags = info.cpu().numpy()
    assert flags[0] != 0 or not np.isfinite(w[0]).all()
    wref = np.linalg.eigvalsh(a)
    for b in range(1, batch):
        assert flags[b] == 0
        assert np.abs(w[b] - wref).max() < 1e-12 * n * max(1.0, np.abs(wref).max())


def test_dgemm_randomised_sweep(env):
    """60 random shapes around the tile edges, all transposes, odd leading dimensions, batches, beta = 0 on NaN-filled C
    (scripts/gemm_fuzz.py)."""
    from scripts.gemm_fuzz import main

    assert main(60, 500) == 0


def test_generalised_eigenproblem_conditioning(env):
    """H C = S C eps through the public solver API for overlaps of condition 3 .. 1e8 (scripts/geneig_fuzz.py): errors
    bounded by what the conditioning allows, C^T S C = I."""
    from scripts.geneig_fuzz import main

    assert main(12, 100) == 0
