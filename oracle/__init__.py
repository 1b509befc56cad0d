"""CPU oracle for the gscf SCF-iteration hot path.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy (LAPACK) restatement of
the algorithm in molsturm/gscf (reference files cited function by function).  It
is imported exclusively by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product path
(``gscf_b200``) never imports it and has no CPU fallback.

Parity pinning status
---------------------
* Pinned: PlainScf / PulayDiisScf / TruncatedOptDampScf on the reference's own
  Be / Sturmian-14 test problem against the golden eigenvalues, eigenvectors
  and energies of ``tests/FunctionalityTest.cc:63-103`` (see
  ``tests/golden/functionality_test.json`` and ``tests/test_oracle_golden.py``).
* The arithmetic of the reference lives in lazyten -> armadillo -> LAPACK,
  which is *not* vendored in the reference tree and cannot be built here
  (no network, unpinned HEAD of github.com/lazyten/{krims,lazyten}).  The oracle
  therefore calls the same LAPACK routine family directly (``dsygvd``,
  ``dgesv``, ``dgemm`` through scipy's OpenBLAS).
* PARITY UNPINNED (no reference test or golden vector exists): the
  unrestricted (alpha/beta block) branch, the damped TODA branch (lambda < 1),
  DIIS ring eviction beyond the 6 iterations of the reference test,
  ``solve_with_guess``, ``n_eigenpairs < n``.  For those the oracle follows the
  reference source line by line but nothing external confirms it.
"""
from .scf import (  # noqa: F401
    ScfResult,
    ScfFailed,
    plain_scf,
    pulay_diis_scf,
    truncated_opt_damp_scf,
    generalised_eigh,
    eig_driver,
    pulay_error,
    compute_damping_coeff,
    diis_linear_system_matrix,
)
