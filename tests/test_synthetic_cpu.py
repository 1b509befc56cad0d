"""The synthetic density-fitted problem of SURVEY.md 8(d): generator known answers (pure-Python integer restatement of
the splitmix64 recipe) and the behaviour the survey probed for it (DIIS converges, plain SCF does not)."""
import numpy as np
import pytest

from oracle import ScfFailed, plain_scf, pulay_diis_scf
from oracle.synthetic import SyntheticDFProblem, splitmix_u, symrand

M64 = (1 << 64) - 1


def u_ref(seed, k):
    z = (seed + (k + 1) * 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return (z >> 11) * 2.0 ** -52 - 1.0


@pytest.mark.parametrize("seed", [0, 1, 1001, 2 ** 40 + 7])
def test_splitmix_known_answers(seed):
    ks = np.array([0, 1, 2, 17, 2 ** 20, 2 ** 33 + 5], dtype=np.uint64)
    got = splitmix_u(seed, ks)
    ref = np.array([u_ref(seed, int(k)) for k in ks])
    assert np.array_equal(got, ref)          # integer recipe: bit-exact
    assert np.all((got >= -1.0) & (got < 1.0))


def test_symrand_is_symmetric_and_indexed_by_the_packed_triangle():
    n, seed = 9, 1002
    a = symrand(seed, n)
    assert np.array_equal(a, a.T)
    for i in range(n):
        for j in range(n):
            hi, lo = max(i, j), min(i, j)
            assert a[i, j] == u_ref(seed, hi * (hi + 1) // 2 + lo)


def test_problem_matrices_follow_the_recipe():
    n, o, aux, p = 24, 3, 4, 5
    prob = SyntheticDFProblem(n, o, n_aux=aux, seed=p)
    S = np.eye(n) + (0.25 / np.sqrt(n)) * symrand(1000 * p + 1, n)
    np.fill_diagonal(S, 1.0)
    assert np.array_equal(prob.S, S)
    h = 2.0 * np.arange(n) / n - 1.0 + 0.5 * (np.arange(n) >= o)
    H = (0.5 / np.sqrt(n)) * symrand(1000 * p + 2, n) + np.diag(h)
    assert np.allclose(prob.H, H, rtol=0, atol=1e-15)
    assert np.linalg.eigvalsh(prob.S).min() > 0.3   # well conditioned overlap (cond ~ 1.8)


def test_survey_probe_behaviour_n128():
    """n = 128, n_occ = 16, n_aux = 16: plain SCF oscillates, DIIS(8) converges in 10 iterations."""
    p = SyntheticDFProblem(128, 16, n_aux=16, seed=0)
    F0, e1, e2 = p.fock(p.core_guess()[None])
    res = pulay_diis_scf(p, F0, e1, e2, n_prev_steps=8)
    assert res.converged and res.n_iter == 10
    with pytest.raises(ScfFailed) as ei:
        plain_scf(p, F0, e1, e2, max_iter=40)
    assert ei.value.kind == "max_iter"
