"""ctypes binding of libgscf_b200.so (the C ABI declared in include/gscf_b200.h).

The library is the product: there is no Python/CPU fallback.  If the shared object is
missing, importing this module raises; if no CUDA device is present every compute entry
point returns GSCF_B200_ERR_CUDA and the wrappers raise ``GscfB200Error``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgscf_b200.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_EIGENSOLVER, ERR_DIIS_STEP, ERR_MAX_ITER, ERR_INVALID_STATE, \
    ERR_CALLBACK, ERR_NOT_IMPLEMENTED = range(9)
HOST, DEVICE = 0, 1
EIG_AUTO, EIG_JACOBI, EIG_TRIDIAG_DC = 0, 1, 2
MAX_HISTORY = 16

EV_BEFORE_ITERATION, EV_NEW_DIAGMAT, EV_UPDATE_EIGENPAIRS, EV_UPDATE_PROBLEM_MATRIX, \
    EV_AFTER_ITERATION = range(5)


class GscfB200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"gscf_b200 status {status}: {msg}")
        self.status = status


class Config(C.Structure):
    _fields_ = [("n_bas", C.c_int), ("n_problems", C.c_int), ("n_blocks", C.c_int),
                ("n_alpha", C.c_int), ("n_beta", C.c_int), ("max_history", C.c_int),
                ("eig_method", C.c_int), ("shared_overlap", C.c_int), ("stream", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("n_eigenpairs", C.c_longlong),
                ("max_error_norm", C.c_double), ("diis_n_prev_steps", C.c_int),
                ("toda_n_prev_steps", C.c_int), ("toda_min_fock_prefactor", C.c_double)]


class FockArgs(C.Structure):
    _fields_ = [("n_bas", C.c_int), ("n_blocks", C.c_int), ("n_alpha", C.c_int), ("n_beta", C.c_int),
                ("n_active", C.c_int), ("active_dev", C.c_void_p), ("active_host", C.POINTER(C.c_int)),
                ("C", C.c_void_p), ("D", C.c_void_p), ("F_out", C.c_void_p), ("ld", C.c_int),
                ("block_stride", C.c_longlong), ("problem_stride", C.c_longlong),
                ("F_problem_stride", C.c_longlong), ("energies_out", C.c_void_p), ("stream", C.c_void_p)]


class Record(C.Structure):
    _fields_ = [("energy_total", C.c_double), ("error_norm", C.c_double), ("n_iter", C.c_int),
                ("converged", C.c_int), ("status", C.c_int), ("reserved", C.c_int)]


FOCK_DEVICE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(FockArgs))
FOCK_HOST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                           C.POINTER(C.c_double))
ERROR_HOST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                            C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))
HOOK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int)

# name -> (restype, argtypes); also the list of symbols include/gscf_b200.h declares
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
_LL = C.c_longlong
SIGNATURES = {
    "gscf_b200_last_error": (C.c_char_p, []),
    "gscf_b200_version_string": (C.c_char_p, []),
    "gscf_b200_launch_count": (C.c_uint64, []),
    "gscf_b200_tensor_map_launch_count": (C.c_uint64, []),
    "gscf_b200_dgemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _P, C.c_int, _LL,
                                  _P, C.c_int, _LL, C.c_double, _P, C.c_int, _LL, C.c_int, _P]),
    "gscf_b200_syev_worksize": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gscf_b200_syev": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _LL, _P, _LL, _P, C.c_int, _LL, C.c_int,
                                 _P, C.c_size_t, _P, _P]),
    "gscf_b200_syevx": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_int, _LL, _P, _LL, _P, C.c_int, _LL, C.c_int,
                                  _P, C.c_size_t, _P, _P]),
    "gscf_b200_frob_multidot_worksize": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "gscf_b200_frob_multidot": (C.c_int, [_P, _P, _LL, _I, C.c_int, C.c_int, C.c_int, C.c_int, _LL, _LL,
                                          C.c_int, _P, C.c_int, _P, C.c_size_t, _P]),
    "gscf_b200_axpby_n": (C.c_int, [_P, _LL, _P, _LL, _LL, _I, C.c_int, _P, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, _P]),
    "gscf_b200_diis_solve": (C.c_int, [_P, C.c_int, _LL, _I, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    "gscf_b200_density": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, _LL, _P, C.c_int, _LL, C.c_int, _P]),
    "gscf_b200_pulay_error_worksize": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "gscf_b200_pulay_error": (C.c_int, [C.c_int, C.c_int, C.c_double, _P, C.c_int, _LL, _P, C.c_int, _LL,
                                        _P, C.c_int, _LL, _P, C.c_int, _LL, _P, _LL, C.c_int, _P]),
    "gscf_b200_copy_matrix": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gscf_b200_compute_damping_coeff": (C.c_double, [C.c_double, C.c_double, C.c_double]),
    "gscf_b200_default_params": (None, [C.POINTER(Params)]),
    "gscf_b200_engine_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "gscf_b200_engine_destroy": (None, [_P]),
    "gscf_b200_engine_stream": (_P, [_P]),
    "gscf_b200_set_overlap": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "gscf_b200_set_fock_builder_device": (C.c_int, [_P, _P, _P]),
    "gscf_b200_set_fock_builder_host": (C.c_int, [_P, FOCK_HOST_FN, _P]),
    "gscf_b200_set_error_host": (C.c_int, [_P, ERROR_HOST_FN, _P]),
    "gscf_b200_set_hook": (C.c_int, [_P, HOOK_FN, _P]),
    "gscf_b200_set_problem_matrix": (C.c_int, [_P, _P, C.c_int, C.c_int, _D, _D]),
    "gscf_b200_set_guess": (C.c_int, [_P, _P, C.c_int, C.c_int, _D]),
    "gscf_b200_set_guess_from_matrix": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "gscf_b200_solve_plain": (C.c_int, [_P, C.POINTER(Params)]),
    "gscf_b200_solve_diis": (C.c_int, [_P, C.POINTER(Params)]),
    "gscf_b200_solve_toda": (C.c_int, [_P, C.POINTER(Params)]),
    "gscf_b200_get_n_iter": (C.c_int, [_P, _I]),
    "gscf_b200_get_converged": (C.c_int, [_P, _I]),
    "gscf_b200_get_status": (C.c_int, [_P, _I]),
    "gscf_b200_get_error_norm": (C.c_int, [_P, _D]),
    "gscf_b200_get_energies": (C.c_int, [_P, _D, _D]),
    "gscf_b200_get_n_mtx_applies": (C.c_int, [_P, C.POINTER(_LL)]),
    "gscf_b200_get_orben": (C.c_int, [_P, _D]),
    "gscf_b200_get_n_eigenpairs": (C.c_int, [_P]),
    "gscf_b200_get_coeff": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_prev_coeff": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_fock": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_diagmat": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_density": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_error": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_history_size": (C.c_int, [_P]),
    "gscf_b200_get_error_entry": (C.c_int, [_P, C.c_int, _D, C.c_int]),
    "gscf_b200_get_fock_entry": (C.c_int, [_P, C.c_int, _D, C.c_int]),
    "gscf_b200_nccl_version": (C.c_int, []),
    "gscf_b200_comm_unique_id": (C.c_int, [_P]),
    "gscf_b200_comm_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "gscf_b200_comm_destroy": (None, [_P]),
    "gscf_b200_ensemble_gather": (C.c_int, [_P, _P, _I, _P]),
    "gscf_b200_get_diis_coefficients": (C.c_int, [_P, _D, _I]),
    "gscf_b200_get_current_diis_coefficients": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_get_error_overlaps": (C.c_int, [_P, C.c_int, C.c_int, _D, _I]),
    "gscf_b200_get_damping_coeff": (C.c_int, [_P, _D]),
    "gscf_b200_get_trace": (C.c_int, [_P, C.c_int, _D, _D, _D, C.c_int, _I]),
    "gscf_b200_get_timings": (C.c_int, [_P, _D, _D]),
    "gscf_b200_get_phase_timings": (C.c_int, [_P, _D, C.c_int]),
    "gscf_b200_set_phase_timing": (C.c_int, [_P, C.c_int]),
    "gscf_b200_dfock_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(_LL), C.c_double, C.POINTER(_P)]),
    "gscf_b200_dfock_destroy": (None, [_P]),
    "gscf_b200_dfock_callback": (_P, []),
    "gscf_b200_dfock_overlap": (_P, [_P]),
    "gscf_b200_dfock_core": (_P, [_P]),
    "gscf_b200_padded_ld": (C.c_int, [C.c_int]),
    "gscf_b200_measure_peaks": (C.c_int, [_D, _D, _D]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m gscf_b200.build` "
            "(or __graft_entry__.build()); gscf_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return (lib.gscf_b200_last_error() or b"").decode()


def check(status: int) -> None:
    if status != OK:
        raise GscfB200Error(status, last_error())
