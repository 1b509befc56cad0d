"""Host-side mirror of gscf's solver / plugin interface on top of the C ABI.

Same names, control-parameter keys, defaults, hooks and error behaviour as the reference
(``src/gscf/{ScfBase,PlainScf,PulayDiisScf,TruncatedOptDampScf}.hh``), so that the parity
tests read like ``tests/FunctionalityTest.cc``.  All arithmetic runs in libgscf_b200.so on the
GPU; this module only marshals buffers and mirrors state.  (The C++ twin of this layer, for
C++ hosts, is include/gscf/*.hh.)
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Dict, List, Optional

import numpy as np

from . import lib as L


# ---------------------------------------------------------------------------------------------
# krims / lazyten surface actually used by gscf
# ---------------------------------------------------------------------------------------------
class GenMap(dict):
    """krims::GenMap subset: string keys, '/'-separated hierarchy, unknown keys ignored."""

    def at(self, key, default=None):
        if key in self:
            return self[key]
        if default is None:
            raise KeyError(key)
        return default

    def exists(self, key) -> bool:
        return key in self

    def submap(self, prefix: str) -> "GenMap":
        pre = prefix.rstrip("/") + "/"
        sub = GenMap({k[len(pre):]: v for k, v in self.items() if k.startswith(pre)})
        nested = self.get(prefix.rstrip("/"))
        if isinstance(nested, dict):  # a map-valued entry, like krims::GenMap::update(key, GenMap)
            sub.update(nested)
        return sub

    def update_key(self, key, value):
        if isinstance(value, GenMap):
            for k, v in value.items():
                self[f"{key}/{k}"] = v
        else:
            self[key] = value


class SolverException(RuntimeError):
    """lazyten::SolverException: carries the failed state (``state.is_failed()``)."""

    def __init__(self, msg: str, state=None):
        super().__init__(msg)
        self.state = state


class ExcInnerEigensolverFailed(SolverException):  # ScfBase.hh:33-36
    pass


class ExcDiisStepFailed(SolverException):  # PulayDiisScf.hh:30-32
    pass


class ExcMaximumNumberOfIterationsReached(SolverException):  # lazyten IterativeWrapper
    pass


class ExcInvalidState(RuntimeError):  # krims::ExcInvalidState
    pass


class OrbitalSpace(enum.Enum):  # FocklikeMatrix_i.hh:28-40
    OCC_ALPHA = 0
    OCC_BETA = 1
    VIRT_ALPHA = 2
    VIRT_BETA = 3


class ScfBaseKeys:  # ScfBaseKeys.cc:24-26 (+ lazyten IterativeWrapperKeys)
    max_iter = "max_iter"
    n_eigenpairs = "n_eigenpairs"
    eigensolver_params = "eigensolver"
    max_error_norm = "max_error_norm"


PlainScfKeys = ScfBaseKeys  # PlainScfKeys.hh:26


class PulayDiisScfKeys(ScfBaseKeys):  # PulayDiisScfKeys.cc:24
    n_prev_steps = "diis_n_prev_steps"


class TruncatedOptDampScfKeys(ScfBaseKeys):  # TruncatedOptDampScfKeys.cc:24-25
    n_prev_steps = "toda_n_prev_steps"
    min_fock_prefactor = "toda_min_fock_prefactor"


ALL = -1  # lazyten::Constants<size_type>::all


# ---------------------------------------------------------------------------------------------
# plugin interface
# ---------------------------------------------------------------------------------------------
class FocklikeMatrix_i:
    """ScfProblemMatrix_i (ScfProblemMatrix_i.hh:32-44) + FocklikeMatrix_i (:53-117).

    Host plugins implement update / energies / indices_orbspace / as_dense in Python; device
    plugins additionally expose ``device_callback()`` -> (function pointer, ctx).
    """

    n_blocks = 1

    def scf_update_key(self) -> str:
        raise NotImplementedError

    def update(self, genmap: Dict[str, np.ndarray]) -> None:
        raise NotImplementedError

    def energy_1e_terms(self) -> float:
        raise NotImplementedError

    def energy_2e_terms(self) -> float:
        raise NotImplementedError

    def indices_orbspace(self, osp: OrbitalSpace):
        raise NotImplementedError

    def n_rows(self) -> int:
        raise NotImplementedError

    def as_dense(self) -> np.ndarray:
        """Current matrix as (n_blocks, n, n) (or (n, n))."""
        raise NotImplementedError

    def device_callback(self):
        return None


class Eigensolution:
    """lazyten::Eigensolution: evalues() ascending, evectors() with vectors as columns."""

    def __init__(self, evalues=None, evectors=None):
        self._w = np.zeros(0) if evalues is None else evalues
        self._v = np.zeros((0, 0)) if evectors is None else evectors

    def evalues(self):
        return self._w

    def evectors(self):
        return self._v

    @property
    def n_vectors(self):
        return 0 if self._v.ndim < 2 else self._v.shape[-1]


class ScfState:
    """ScfStateBase (ScfStateBase.hh:54-258) + the public members of PulayDiisScfState
    (PulayDiisScf.hh:38-94) and TruncatedOptDampScfState (TruncatedOptDampScf.hh:39-77)."""

    def __init__(self, problem_matrix, overlap):
        self.problem_matrix_ptr = problem_matrix
        self._overlap = overlap
        self.last_error_norm = float("nan")  # Constants<real>::invalid
        self._n_iter = 0
        self._n_mtx_applies = 0
        self._failed = False
        self.fail_reason = ""
        self._esol = Eigensolution()
        self._prev_esol = Eigensolution()
        self.diis_coefficients = np.zeros(0)
        self.error_overlaps: List[List[float]] = []
        self.errors_back = None
        self.damping_coeff = 1.0
        self.energy_1e = float("nan")
        self.energy_2e = float("nan")
        self.trace_energy: List[float] = []
        self.trace_error: List[float] = []
        self.trace_coeff: List[np.ndarray] = []
        self.scf_ms = 0.0
        self.fock_ms = 0.0
        self.phase_ms = None
        self._engine = None  # live only while hooks run

    # lazyten IterativeStateWrapper / SolverStateBase
    def n_iter(self):
        return self._n_iter

    def n_mtx_applies(self):
        return self._n_mtx_applies

    def is_failed(self):
        return self._failed

    def fail(self, reason):
        self._failed = True
        self.fail_reason = reason

    def problem_matrix(self):
        return self.problem_matrix_ptr

    def overlap_matrix(self):
        return self._overlap

    def eigensolution(self):
        if self._engine is not None:
            self._engine.refresh_eigensolution(self)
        return self._esol

    def previous_eigensolution(self):
        return self._prev_esol

    def diagonalised_matrix(self):
        if self._engine is None:
            raise ExcInvalidState("diagonalised_matrix is only available inside hooks")
        return self._engine.get_matrix("diagmat")

    @property
    def energy_total(self):
        return self.energy_1e + self.energy_2e


# ---------------------------------------------------------------------------------------------
# engine wrapper
# ---------------------------------------------------------------------------------------------
def _dense_blocks(F, n):
    F = np.asarray(F, dtype=np.float64)
    if F.ndim == 2:
        F = F[None]
    assert F.shape[1:] == (n, n), F.shape
    return F


def _to_colmajor(mats: np.ndarray) -> np.ndarray:
    """(k, n, n) row-indexed numpy -> flat buffer of k column-major n x n matrices (ld = n)."""
    return np.ascontiguousarray(np.transpose(mats, (0, 2, 1))).ravel()


def _from_colmajor(buf: np.ndarray, k: int, n: int) -> np.ndarray:
    return np.transpose(buf.reshape(k, n, n), (0, 2, 1)).copy()


class Engine:
    """RAII wrapper of gscf_b200_engine for one or many problems."""

    def __init__(self, n_bas, n_alpha, n_beta=None, n_blocks=1, n_problems=1, max_history=4,
                 eig_method=L.EIG_AUTO, shared_overlap=False):
        n_beta = n_alpha if n_beta is None else n_beta
        self.cfg = L.Config(n_bas, n_problems, n_blocks, n_alpha, n_beta, max_history, eig_method,
                            1 if shared_overlap else 0, None)
        self.n, self.P, self.nblk = n_bas, n_problems, n_blocks
        self.handle = C.c_void_p()
        L.check(L.lib.gscf_b200_engine_create(C.byref(self.cfg), C.byref(self.handle)))
        self._keep = []  # keep ctypes callbacks alive

    def close(self):
        if self.handle:
            L.lib.gscf_b200_engine_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup ---------------------------------------------------------------------------------
    def set_overlap_host(self, S):
        S = np.asarray(S, dtype=np.float64)
        S = S[None] if S.ndim == 2 else S
        buf = _to_colmajor(S)
        L.check(L.lib.gscf_b200_set_overlap(self.handle, buf.ctypes.data, self.n, L.HOST))

    def set_overlap_device(self, ptr, ld):
        L.check(L.lib.gscf_b200_set_overlap(self.handle, ptr, ld, L.DEVICE))

    def set_problem_matrix_host(self, F, e1, e2):
        F = np.asarray(F, dtype=np.float64).reshape(self.P * self.nblk, self.n, self.n)
        buf = _to_colmajor(F)
        e1 = np.ascontiguousarray(np.atleast_1d(e1), dtype=np.float64)
        e2 = np.ascontiguousarray(np.atleast_1d(e2), dtype=np.float64)
        L.check(L.lib.gscf_b200_set_problem_matrix(
            self.handle, buf.ctypes.data, self.n, L.HOST,
            e1.ctypes.data_as(C.POINTER(C.c_double)), e2.ctypes.data_as(C.POINTER(C.c_double))))

    def set_problem_matrix_device(self, ptr, ld, e1=None, e2=None):
        p1 = None if e1 is None else np.ascontiguousarray(e1, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        p2 = None if e2 is None else np.ascontiguousarray(e2, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        L.check(L.lib.gscf_b200_set_problem_matrix(self.handle, ptr, ld, L.DEVICE, p1, p2))

    def set_guess_host(self, Cg, orben=None):
        Cg = np.asarray(Cg, dtype=np.float64).reshape(self.P * self.nblk, self.n, self.n)
        buf = _to_colmajor(Cg)
        ob = None
        if orben is not None:
            ob = np.ascontiguousarray(orben, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        L.check(L.lib.gscf_b200_set_guess(self.handle, buf.ctypes.data, self.n, L.HOST, ob))

    def set_fock_device(self, fn_ptr, ctx):
        L.check(L.lib.gscf_b200_set_fock_builder_device(self.handle, fn_ptr, ctx))

    def set_fock_host(self, pyfn):
        cb = L.FOCK_HOST_FN(pyfn)
        self._keep.append(cb)
        L.check(L.lib.gscf_b200_set_fock_builder_host(self.handle, cb, None))

    def set_error_host(self, pyfn):
        cb = L.ERROR_HOST_FN(pyfn)
        self._keep.append(cb)
        L.check(L.lib.gscf_b200_set_error_host(self.handle, cb, None))

    def set_hook(self, pyfn):
        cb = L.HOOK_FN(pyfn)
        self._keep.append(cb)
        L.check(L.lib.gscf_b200_set_hook(self.handle, cb, None))

    def set_phase_timing(self, on=True):
        L.check(L.lib.gscf_b200_set_phase_timing(self.handle, 1 if on else 0))

    # -- solve ---------------------------------------------------------------------------------
    def solve(self, method: str, params: L.Params) -> int:
        fn = {"plain": L.lib.gscf_b200_solve_plain, "diis": L.lib.gscf_b200_solve_diis,
              "toda": L.lib.gscf_b200_solve_toda}[method]
        return fn(self.handle, C.byref(params))

    # -- results -------------------------------------------------------------------------------
    def _ivec(self, fn):
        out = np.zeros(self.P, dtype=np.int32)
        L.check(fn(self.handle, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def _dvec(self, fn):
        out = np.zeros(self.P, dtype=np.float64)
        L.check(fn(self.handle, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def n_iter(self):
        return self._ivec(L.lib.gscf_b200_get_n_iter)

    def converged(self):
        return self._ivec(L.lib.gscf_b200_get_converged)

    def status(self):
        return self._ivec(L.lib.gscf_b200_get_status)

    def error_norm(self):
        return self._dvec(L.lib.gscf_b200_get_error_norm)

    def damping_coeff(self):
        return self._dvec(L.lib.gscf_b200_get_damping_coeff)

    def energies(self):
        e1 = np.zeros(self.P)
        e2 = np.zeros(self.P)
        L.check(L.lib.gscf_b200_get_energies(self.handle, e1.ctypes.data_as(C.POINTER(C.c_double)),
                                             e2.ctypes.data_as(C.POINTER(C.c_double))))
        return e1, e2

    def orben(self):
        out = np.zeros((self.P, self.nblk, self.n))
        L.check(L.lib.gscf_b200_get_orben(self.handle, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def get_matrix(self, which: str):
        fn = {"coeff": L.lib.gscf_b200_get_coeff, "prev_coeff": L.lib.gscf_b200_get_prev_coeff,
              "fock": L.lib.gscf_b200_get_fock, "diagmat": L.lib.gscf_b200_get_diagmat,
              "density": L.lib.gscf_b200_get_density, "error": L.lib.gscf_b200_get_error}[which]
        buf = np.zeros(self.P * self.nblk * self.n * self.n)
        L.check(fn(self.handle, buf.ctypes.data_as(C.POINTER(C.c_double)), self.n))
        return _from_colmajor(buf, self.P * self.nblk, self.n).reshape(self.P, self.nblk, self.n, self.n)

    def diis_coefficients(self):
        out = np.zeros((self.P, L.MAX_HISTORY))
        k = C.c_int(0)
        L.check(L.lib.gscf_b200_get_diis_coefficients(self.handle, out.ctypes.data_as(C.POINTER(C.c_double)),
                                                      C.byref(k)))
        return out[:, : k.value]

    def current_diis_coefficients(self, problem=0):
        """Coefficients of the extrapolation being diagonalised right now (valid inside the NEW_DIAGMAT /
        UPDATE_EIGENPAIRS hooks): [] before the first error, [1] with one, the device solution otherwise."""
        k = int(L.lib.gscf_b200_get_history_size(self.handle))
        if k <= 0:
            return np.zeros(0)
        if k == 1:
            return np.ones(1)
        out = np.zeros((self.P, k))
        L.check(L.lib.gscf_b200_get_current_diis_coefficients(self.handle, out.ctypes.data_as(C.POINTER(C.c_double)), k))
        return out[problem]

    def error_overlaps(self, problem=0):
        res = []
        i = 0
        while True:
            ov = np.zeros(L.MAX_HISTORY)
            ln = C.c_int(0)
            st = L.lib.gscf_b200_get_error_overlaps(self.handle, problem, i,
                                                    ov.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ln))
            if st != L.OK:
                break
            res.append(list(ov[: ln.value]))
            i += 1
        return res

    def trace(self, problem=0):
        rows = 512
        en = np.zeros(rows)
        er = np.zeros(rows)
        cf = np.zeros((rows, L.MAX_HISTORY))
        n = C.c_int(0)
        L.check(L.lib.gscf_b200_get_trace(self.handle, problem, en.ctypes.data_as(C.POINTER(C.c_double)),
                                          er.ctypes.data_as(C.POINTER(C.c_double)),
                                          cf.ctypes.data_as(C.POINTER(C.c_double)), rows, C.byref(n)))
        k = min(n.value, rows)
        coeffs = [row[~np.isnan(row)] for row in cf[:k]]
        return list(en[:k]), list(er[:k]), coeffs

    def timings(self):
        a = C.c_double(0)
        b = C.c_double(0)
        L.check(L.lib.gscf_b200_get_timings(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def phase_timings(self):
        out = np.zeros(32)
        L.check(L.lib.gscf_b200_get_phase_timings(self.handle, out.ctypes.data_as(C.POINTER(C.c_double)), 32))
        names = {0: "diagmat", 1: "eig", 2: "density", 3: "error", 4: "coeff", 5: "fock", 6: "total", 7: "overlap",
                 8: "tridiag", 9: "stedc", 10: "backtransform", 11: "panel_kernel", 12: "ortho_gemm",
                 13: "trailing_gemm", 14: "guess_fock"}
        res = {v: float(out[k]) for k, v in names.items()}
        res["counts"] = {v: int(out[16 + k]) for k, v in names.items()}
        return res

    # used by ScfState while hooks run
    def n_eigenpairs(self) -> int:
        """Eigenpairs kept by the last solve (ScfBaseKeys.n_eigenpairs, ScfBase.hh:87)."""
        return int(L.lib.gscf_b200_get_n_eigenpairs(self.handle))

    def refresh_eigensolution(self, state):
        k = self.n_eigenpairs()
        w = self.orben()[0][..., :k]
        Cm = self.get_matrix("coeff")[0][..., :k]
        if self.nblk == 1:
            state._esol = Eigensolution(w[0], Cm[0])
        else:
            state._esol = Eigensolution(w, Cm)


# ---------------------------------------------------------------------------------------------
# the solver classes
# ---------------------------------------------------------------------------------------------
class ScfBase:
    """gscf::ScfBase (ScfBase.hh:48-262)."""

    _method = None

    def __init__(self, params: Optional[dict] = None):
        self.max_iter = 100  # lazyten IterativeWrapper default
        self.n_eigenpairs = ALL  # ScfBase.hh:87
        self.eigensolver_params = GenMap()  # :90
        self.max_error_norm = 5e-7  # :94
        self.user_eigensolver_tolerance = False  # :109
        self.eig_method = L.EIG_AUTO  # "b200/eig_method": new keys live under a new prefix
        if params:
            self.update_control_params(GenMap(params))

    # -- control parameters (ScfBase.hh:100-119) -----------------------------------------------
    def update_control_params(self, m: GenMap):
        m = m if isinstance(m, GenMap) else GenMap(m)
        self.max_iter = m.at(ScfBaseKeys.max_iter, self.max_iter)
        self.n_eigenpairs = m.at(ScfBaseKeys.n_eigenpairs, self.n_eigenpairs)
        self.max_error_norm = m.at(ScfBaseKeys.max_error_norm, self.max_error_norm)
        self.eigensolver_params = m.submap(ScfBaseKeys.eigensolver_params)
        self.user_eigensolver_tolerance = m.exists("tolerance")
        self.eig_method = m.at("b200/eig_method", self.eig_method)
        # The "eigensolver" sub-map goes to the inner eigensolver (ScfBase.hh:104, :324): lazyten's key "method" picks
        # among solvers that return the same eigenpairs; here it picks among this library's dense solvers.  Values lazyten
        # knows ("auto", "lapack", "arpack") all map to the automatic choice, plus "jacobi" and "tridiag_dc"; its
        # "tolerance" has no meaning for a direct solver (inner_solver_tolerance is machine epsilon anyway, ScfBase.hh:251).
        method = self.eigensolver_params.get("method")
        if isinstance(method, str):
            self.eig_method = {"jacobi": L.EIG_JACOBI, "tridiag_dc": L.EIG_TRIDIAG_DC, "dc": L.EIG_TRIDIAG_DC}.get(
                method.lower(), L.EIG_AUTO)

    def get_control_params(self, m: GenMap):
        m.update_key(ScfBaseKeys.max_iter, self.max_iter)
        m.update_key(ScfBaseKeys.n_eigenpairs, self.n_eigenpairs)
        m.update_key(ScfBaseKeys.max_error_norm, self.max_error_norm)
        m.update_key(ScfBaseKeys.eigensolver_params, self.eigensolver_params)

    def is_converged(self, s: ScfState) -> bool:  # ScfBase.hh:123-126
        return s.last_error_norm < self.max_error_norm

    # -- hooks (ScfBase.hh:195,200; lazyten before/after_iteration_step) --------------------------
    def before_iteration_step(self, s):
        pass

    def after_iteration_step(self, s):
        pass

    def on_update_eigenpairs(self, s):
        pass

    def on_update_problem_matrix(self, s):
        pass

    def on_new_diagmat(self, s):
        pass

    # Override to replace the built-in device Pulay error (ScfBase.hh:217).  Signature of the
    # override: calculate_error(state, F, C, orben) -> (n_blocks, n, n) array.
    calculate_error = None

    # -- C-struct parameters ------------------------------------------------------------------------
    def _params(self) -> L.Params:
        p = L.Params()
        L.lib.gscf_b200_default_params(C.byref(p))
        p.max_iter = int(self.max_iter)
        p.n_eigenpairs = int(self.n_eigenpairs)
        p.max_error_norm = float(self.max_error_norm)
        return p

    def _history(self) -> int:
        return 2

    # -- run ------------------------------------------------------------------------------------------
    def solve(self, probmat, overlap) -> ScfState:  # ScfBase.hh:140-145
        state = ScfState(probmat, overlap)
        self.solve_state(state)
        return state

    def solve_with_guess(self, probmat, overlap, guess_state: ScfState) -> ScfState:
        """ScfBase.hh:152-163: install the guess eigensolution, rebuild F, then iterate."""
        state = ScfState(probmat, overlap)
        state._prev_esol = guess_state.previous_eigensolution()
        state._guess = guess_state.eigensolution()
        self.solve_state(state)
        return state

    def solve_state(self, state: ScfState):
        if state.is_failed():
            raise ExcInvalidState("Cannot solve a failed state")
        pm = state.problem_matrix_ptr
        n = pm.n_rows()
        nblk = getattr(pm, "n_blocks", 1)
        oa = pm.indices_orbspace(OrbitalSpace.OCC_ALPHA)
        ob = pm.indices_orbspace(OrbitalSpace.OCC_BETA)
        n_alpha = oa[1] - oa[0]
        n_beta = ob[1] - ob[0]
        eng = Engine(n, n_alpha, n_beta, n_blocks=nblk, n_problems=1, max_history=self._history(),
                     eig_method=self.eig_method)
        try:
            eng.set_overlap_host(state.overlap_matrix())
            dev = pm.device_callback()
            if dev is not None:
                eng.set_fock_device(dev[0], dev[1])
            else:
                key = pm.scf_update_key()

                def fock_cb(ctx, problem, nb, nblocks, Cptr, wptr, Fptr, e1ptr, e2ptr):
                    try:
                        Cb = np.ctypeslib.as_array(Cptr, shape=(nblocks * nb * nb,))
                        Cm = _from_colmajor(Cb, nblocks, nb)
                        pm.update({key: Cm[0] if nblocks == 1 else Cm})
                        F = _dense_blocks(pm.as_dense(), nb)
                        Fb = np.ctypeslib.as_array(Fptr, shape=(nblocks * nb * nb,))
                        Fb[:] = _to_colmajor(F)
                        e1ptr[0] = pm.energy_1e_terms()
                        e2ptr[0] = pm.energy_2e_terms()
                        return 0
                    except Exception as exc:  # pragma: no cover - surfaced as ERR_CALLBACK
                        state.fail_reason = repr(exc)
                        return 1

                eng.set_fock_host(fock_cb)
            if type(self).calculate_error is not None:
                def err_cb(ctx, problem, nb, nblocks, Fptr, Cptr, wptr, Eptr):
                    try:
                        Fm = _from_colmajor(np.ctypeslib.as_array(Fptr, shape=(nblocks * nb * nb,)), nblocks, nb)
                        Cm = _from_colmajor(np.ctypeslib.as_array(Cptr, shape=(nblocks * nb * nb,)), nblocks, nb)
                        w = np.ctypeslib.as_array(wptr, shape=(nblocks * nb,)).reshape(nblocks, nb).copy()
                        E = _dense_blocks(self.calculate_error(state, Fm, Cm, w), nb)
                        np.ctypeslib.as_array(Eptr, shape=(nblocks * nb * nb,))[:] = _to_colmajor(E)
                        return 0
                    except Exception as exc:  # pragma: no cover
                        state.fail_reason = repr(exc)
                        return 1

                eng.set_error_host(err_cb)
            guess = getattr(state, "_guess", None)
            if guess is not None:  # obtain_guess_from, ScfStateBase.hh:291-303
                Cg = np.asarray(guess.evectors(), dtype=np.float64)
                Cfull = np.zeros((nblk, n, n))
                Cg = Cg[None] if Cg.ndim == 2 else Cg
                Cfull[:, :, : Cg.shape[-1]] = Cg
                eng.set_guess_host(Cfull)
            else:
                eng.set_problem_matrix_host(_dense_blocks(pm.as_dense(), n), [pm.energy_1e_terms()],
                                            [pm.energy_2e_terms()])
            hooks = {L.EV_BEFORE_ITERATION: "before_iteration_step", L.EV_NEW_DIAGMAT: "on_new_diagmat",
                     L.EV_UPDATE_EIGENPAIRS: "on_update_eigenpairs",
                     L.EV_UPDATE_PROBLEM_MATRIX: "on_update_problem_matrix",
                     L.EV_AFTER_ITERATION: "after_iteration_step"}
            overridden = any(getattr(type(self), h) is not getattr(ScfBase, h) for h in hooks.values())
            if overridden:
                def hook_cb(ctx, ev, it):
                    try:
                        state._n_iter = it
                        state._engine = eng
                        if ev in (L.EV_UPDATE_PROBLEM_MATRIX, L.EV_AFTER_ITERATION):
                            self._mirror_scalars(eng, state)
                        if self._method == "diis" and ev == L.EV_NEW_DIAGMAT:
                            # public state member (PulayDiisScf.hh:83-86): the coefficients of THIS iteration's
                            # extrapolation live on the device until the iteration's scalars are fetched
                            state.diis_coefficients = eng.current_diis_coefficients()
                        elif self._method == "diis" and ev == L.EV_AFTER_ITERATION:
                            c = eng.diis_coefficients()  # the row recorded for the iteration that just ended
                            state.diis_coefficients = c[0] if c.size else np.zeros(0)
                        getattr(self, hooks[ev])(state)
                        return 0
                    except Exception as exc:  # pragma: no cover
                        state.fail_reason = repr(exc)
                        return 1
                    finally:
                        state._engine = None

                eng.set_hook(hook_cb)
            st = eng.solve(self._method, self._params())
            self._mirror(eng, state)
            if st == L.ERR_EIGENSOLVER:
                self._rescue_failed_eigenproblem(eng, state)
            if st != L.OK:
                msg = L.last_error()
                state.fail(msg)
                exc = {L.ERR_MAX_ITER: ExcMaximumNumberOfIterationsReached,
                       L.ERR_DIIS_STEP: ExcDiisStepFailed,
                       L.ERR_EIGENSOLVER: ExcInnerEigensolverFailed}.get(st)
                if exc is None:
                    raise L.GscfB200Error(st, msg + " " + state.fail_reason)
                raise exc(msg, state)
        finally:
            eng.close()
        return state

    def _rescue_failed_eigenproblem(self, eng: Engine, state: ScfState):
        """lazyten::rescue::failed_eigenproblem (ScfBase.hh:334): when the inner eigensolver fails, the eigenproblem that
        failed is dumped for a post-mortem - here the matrix that was being diagonalised and the overlap, as
        ``failed_eigenproblem_<pid>.npz`` in the directory named by GSCF_B200_RESCUE_DIR (no dump when unset)."""
        import os

        d = os.environ.get("GSCF_B200_RESCUE_DIR")
        if not d:
            return
        try:
            os.makedirs(d, exist_ok=True)
            path = os.path.join(d, f"failed_eigenproblem_{os.getpid()}.npz")
            np.savez(path, diagonalised_matrix=eng.get_matrix("diagmat")[0], overlap=np.asarray(state.overlap_matrix()),
                     n_iter=int(eng.n_iter()[0]))
            state.rescue_path = path
        except Exception:  # noqa: BLE001 - a failed dump must not mask the solver exception
            pass

    def _mirror_scalars(self, eng: Engine, state: ScfState):
        e1, e2 = eng.energies()
        state.energy_1e, state.energy_2e = float(e1[0]), float(e2[0])
        state.last_error_norm = float(eng.error_norm()[0])
        state.damping_coeff = float(eng.damping_coeff()[0])

    def _mirror(self, eng: Engine, state: ScfState):
        state._n_iter = int(eng.n_iter()[0])
        self._mirror_scalars(eng, state)
        if state._n_iter > 0:
            state._prev_esol = Eigensolution(*self._esol_arrays(eng, "prev_coeff", None))
            eng.refresh_eigensolution(state)
            state.errors_back = eng.get_matrix("error")[0]
        state.trace_energy, state.trace_error, state.trace_coeff = eng.trace(0)
        state.scf_ms, state.fock_ms = eng.timings()
        state.phase_ms = eng.phase_timings()
        state.density = eng.get_matrix("density")[0] if state._n_iter > 0 else None
        state.fock = eng.get_matrix("fock")[0]
        if self._method == "diis":
            c = eng.diis_coefficients()
            state.diis_coefficients = c[0] if c.size else np.zeros(0)
            state.error_overlaps = eng.error_overlaps(0)

    def _esol_arrays(self, eng, which, w):
        Cm = eng.get_matrix(which)[0]
        return (w, Cm[0] if eng.nblk == 1 else Cm)


class PlainScf(ScfBase):  # PlainScf.hh:71-147
    _method = "plain"


class PulayDiisScf(ScfBase):  # PulayDiisScf.hh:139-256
    _method = "diis"

    def __init__(self, params: Optional[dict] = None):
        self.n_prev_steps = 4  # PulayDiisScf.hh:171
        super().__init__(params)

    def update_control_params(self, m):
        m = m if isinstance(m, GenMap) else GenMap(m)
        super().update_control_params(m)
        self.n_prev_steps = m.at(PulayDiisScfKeys.n_prev_steps, self.n_prev_steps)

    def get_control_params(self, m):
        super().get_control_params(m)
        m.update_key(PulayDiisScfKeys.n_prev_steps, self.n_prev_steps)

    def _params(self):
        p = super()._params()
        p.diis_n_prev_steps = int(self.n_prev_steps)
        return p

    def _history(self):
        return int(self.n_prev_steps)


class TruncatedOptDampScf(ScfBase):  # TruncatedOptDampScf.hh:124-233
    _method = "toda"

    def __init__(self, params: Optional[dict] = None):
        self.n_prev_steps = 2  # TruncatedOptDampScf.hh:159
        self.min_fock_prefactor = 0.05  # :166
        super().__init__(params)

    def update_control_params(self, m):
        m = m if isinstance(m, GenMap) else GenMap(m)
        super().update_control_params(m)
        self.n_prev_steps = m.at(TruncatedOptDampScfKeys.n_prev_steps, self.n_prev_steps)
        self.min_fock_prefactor = m.at(TruncatedOptDampScfKeys.min_fock_prefactor, self.min_fock_prefactor)

    def get_control_params(self, m):
        super().get_control_params(m)
        m.update_key(TruncatedOptDampScfKeys.n_prev_steps, self.n_prev_steps)
        m.update_key(TruncatedOptDampScfKeys.min_fock_prefactor, self.min_fock_prefactor)

    def _params(self):
        p = super()._params()
        p.toda_n_prev_steps = int(self.n_prev_steps)
        p.toda_min_fock_prefactor = float(self.min_fock_prefactor)
        return p


# ---------------------------------------------------------------------------------------------
# device plugin: synthetic density-fitted Fock builder
# ---------------------------------------------------------------------------------------------
class DeviceDFFock(FocklikeMatrix_i):
    """Device-resident synthetic DF problem matrices for P problems (SURVEY 8d generator)."""

    def __init__(self, n_bas, n_alpha, n_beta=None, n_aux=None, seeds=(0,), g=1.0, unrestricted=None):
        n_beta = n_alpha if n_beta is None else n_beta
        unrestricted = (n_alpha != n_beta) if unrestricted is None else unrestricted
        self.n_bas, self.n_alpha, self.n_beta = n_bas, n_alpha, n_beta
        self.n_blocks = 2 if unrestricted else 1
        self.n_aux = n_aux if n_aux is not None else max(1, n_bas // 16)
        self.seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        self.P = len(self.seeds)
        self.handle = C.c_void_p()
        L.check(L.lib.gscf_b200_dfock_create(n_bas, n_alpha, n_beta, self.n_blocks, self.n_aux, self.P,
                                             self.seeds.ctypes.data_as(C.POINTER(C.c_longlong)), g,
                                             C.byref(self.handle)))
        self.ld = L.lib.gscf_b200_padded_ld(n_bas)

    def close(self):
        if self.handle:
            L.lib.gscf_b200_dfock_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def scf_update_key(self):
        return "evec_coefficients"

    def n_rows(self):
        return self.n_bas

    def indices_orbspace(self, osp):
        n = self.n_bas
        return {OrbitalSpace.OCC_ALPHA: (0, self.n_alpha), OrbitalSpace.OCC_BETA: (0, self.n_beta)
                if self.n_blocks == 1 else (n, n + self.n_beta),
                OrbitalSpace.VIRT_ALPHA: (self.n_alpha, n),
                OrbitalSpace.VIRT_BETA: (self.n_beta, n) if self.n_blocks == 1 else (n + self.n_beta, 2 * n)}[osp]

    def device_callback(self):
        return L.lib.gscf_b200_dfock_callback(), self.handle

    def overlap_ptr(self):
        return L.lib.gscf_b200_dfock_overlap(self.handle)

    def core_ptr(self):
        return L.lib.gscf_b200_dfock_core(self.handle)


def solve_device_problem(fock: DeviceDFFock, method="diis", params: Optional[dict] = None,
                         max_history=None, eig_method=L.EIG_AUTO, phase_timing=False):
    """Whole-loop, device-resident solve of the P problems of a DeviceDFFock plugin, starting
    from the core guess (eigenvectors of H against S).  Returns the live Engine (caller closes).
    """
    params = GenMap(params or {})
    cls = {"plain": PlainScf, "diis": PulayDiisScf, "toda": TruncatedOptDampScf}[method]
    solver = cls(params)
    hist = max_history if max_history is not None else max(2, solver._history())
    eng = Engine(fock.n_bas, fock.n_alpha, fock.n_beta, n_blocks=fock.n_blocks, n_problems=fock.P,
                 max_history=hist, eig_method=eig_method)
    try:
        eng.set_overlap_device(fock.overlap_ptr(), fock.ld)
        fn, ctx = fock.device_callback()
        eng.set_fock_device(fn, ctx)
        eng.set_phase_timing(phase_timing)
        core_guess_device(eng, fock)
        status = eng.solve(method, solver._params())
    except Exception:
        eng.close()
        raise
    return eng, status


def core_guess_device(eng: Engine, fock: DeviceDFFock):
    """Core guess: eigenvectors of H against S, then F(C_guess) through the plugin
    (``solve_with_guess`` semantics, ScfStateBase.hh:291-303)."""
    L.check(L.lib.gscf_b200_set_guess_from_matrix(eng.handle, fock.core_ptr(), fock.ld, L.DEVICE))
