"""Observability helpers mirroring the reference example.

* ``MathematicaWriter`` writes scalars / vectors / matrices as ``name = {{...},{...}};`` lines - the format
  ``lazyten::io::make_writer<io::Mathematica>`` produces for ``examples/hartree_fock/main.cc:49-54``
  (``/tmp/debug_gscf_*.m``).
* ``verbose(SolverClass, writer, out)`` returns a solver class whose hooks print and dump per iteration what
  ``examples/hartree_fock/ScfLibrary.hh:33-103`` (``VerboseScf``) does: orbital energies and coefficients after the
  eigensolve, energies / Fock matrix / error after the problem-matrix update, the Pulay error norm at the end of the step,
  the extrapolated matrix on ``on_new_diagmat`` (``ScfLibrary.hh:114-133``).
* ``IterationRecorder`` keeps the same per-iteration records in memory and can save them as ``.npz`` / JSON, so that a
  trace of this implementation can be diffed against a trace of the oracle.

Hooks force a device -> host copy of what they touch; they are meant for debugging, not for timed runs.
"""
from __future__ import annotations

import io
import json
import sys
from typing import Dict, List, Optional

import numpy as np


def _fmt(x: float, precision: int) -> str:
    s = np.format_float_scientific(float(x), precision=precision, unique=False, exp_digits=1)
    mant, exp = s.split("e")
    return mant if int(exp) == 0 else f"{mant}*^{int(exp)}"  # Mathematica: 1.5*^-7


class MathematicaWriter:
    """``DataWriter_i``: write(name, value) for scalars, vectors and matrices."""

    def __init__(self, stream=None, precision: int = 15):
        self.stream = stream if stream is not None else io.StringIO()
        self.precision = precision

    def write(self, name: str, value) -> None:
        a = np.asarray(value, dtype=np.float64)
        if a.ndim == 0:
            body = _fmt(a, self.precision)
        elif a.ndim == 1:
            body = "{" + ",".join(_fmt(v, self.precision) for v in a) + "}"
        elif a.ndim == 2:
            body = "{" + ",".join("{" + ",".join(_fmt(v, self.precision) for v in row) + "}" for row in a) + "}"
        else:
            raise ValueError("MathematicaWriter: at most two dimensions")
        self.stream.write(f"{name} = {body};\n")

    def write_comment(self, text: str) -> None:
        self.stream.write(f"(* {text} *)\n")

    def getvalue(self) -> str:
        return self.stream.getvalue() if hasattr(self.stream, "getvalue") else ""


class IterationRecorder:
    """Per-iteration records {iter, energy_1e, energy_2e, energy_total, error_norm, orben, coefficients}."""

    def __init__(self, keep_matrices: bool = False):
        self.keep_matrices = keep_matrices
        self.records: List[Dict] = []

    def record(self, state, **extra) -> None:
        rec = {"iter": int(state.n_iter()), "energy_1e": float(state.energy_1e), "energy_2e": float(state.energy_2e),
               "energy_total": float(state.energy_total), "error_norm": float(state.last_error_norm)}
        if getattr(state, "diis_coefficients", None) is not None and len(state.diis_coefficients):
            rec["diis_coefficients"] = np.asarray(state.diis_coefficients, dtype=float).tolist()
        rec["damping_coeff"] = float(getattr(state, "damping_coeff", 1.0))
        rec.update(extra)
        self.records.append(rec)

    def to_json(self, path: Optional[str] = None) -> str:
        def clean(v):
            return v.tolist() if isinstance(v, np.ndarray) else v

        text = json.dumps([{k: clean(v) for k, v in r.items()} for r in self.records], indent=1)
        if path:
            with open(path, "w") as f:
                f.write(text)
        return text

    def to_npz(self, path: str) -> None:
        cols = {}
        for key in ("iter", "energy_1e", "energy_2e", "energy_total", "error_norm", "damping_coeff"):
            cols[key] = np.array([r[key] for r in self.records])
        for r in self.records:
            for k, v in r.items():
                if isinstance(v, np.ndarray):
                    cols[f"{k}_{r['iter']}"] = v
        np.savez(path, **cols)


def verbose(solver_cls, writer: Optional[MathematicaWriter] = None, out=None, recorder: Optional[IterationRecorder] = None):
    """``VerboseScf<Solver>`` (ScfLibrary.hh:33-103): subclass of ``solver_cls`` with printing / dumping hooks."""
    out = out if out is not None else sys.stdout

    class VerboseScf(solver_cls):  # type: ignore[misc, valid-type]
        def before_iteration_step(self, s):
            out.write(f"\nStarting SCF iteration {s.n_iter()}\n")
            super().before_iteration_step(s)

        def on_update_eigenpairs(self, s):
            es = s.eigensolution()
            ev = np.atleast_2d(np.asarray(es.evalues()))
            out.write("   New orbital eigenvalues: \n        " + " ".join(f"{v:.6g}" for v in ev.ravel()) + " \n")
            if writer is not None:
                writer.write(f"evals{s.n_iter()}", ev[0] if ev.shape[0] == 1 else ev)
                C = np.asarray(es.evectors())
                writer.write(f"evecs{s.n_iter()}", C if C.ndim == 2 else C.reshape(-1, C.shape[-1]))
            super().on_update_eigenpairs(s)

        def on_new_diagmat(self, s):
            if writer is not None:
                D = np.asarray(s.diagonalised_matrix())
                writer.write(f"diagmat{s.n_iter()}", D.reshape(-1, D.shape[-1]))
            super().on_new_diagmat(s)

        def on_update_problem_matrix(self, s):
            out.write("   Current HF energies:\n"
                      f"     E_1e     = {s.energy_1e}\n     E_2e     = {s.energy_2e}\n\n"
                      f"     E_total  = {s.energy_total:.15g}\n")
            super().on_update_problem_matrix(s)

        def after_iteration_step(self, s):
            out.write(f"   Current Pulay error: {s.last_error_norm}\n")
            if writer is not None:
                writer.write(f"pulayerror{s.n_iter()}", s.last_error_norm)
            if recorder is not None:
                recorder.record(s)
            super().after_iteration_step(s)

    VerboseScf.__name__ = f"VerboseScf[{solver_cls.__name__}]"
    return VerboseScf
