"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/gscf_b200.h declares; without a GPU the compute entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "gscf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gscf_b200_[a-z0-9_]+)\s*\(", src))
    # function-pointer typedefs are not exported symbols
    names -= set(re.findall(r"\(\*\s*(gscf_b200_[a-z0-9_]+)\s*\)", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from gscf_b200 import lib as L

    raw = C.CDLL(L.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) > 40
    missing = [s for s in declared if not hasattr(raw, s)]
    assert not missing, missing
    # and the ctypes table covers the header
    assert set(declared) == set(L.SIGNATURES.keys())


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gscf_b200 import lib as L

    h = C.c_void_p()
    cfg = L.Config(14, 1, 1, 2, 2, 4, 0, 0, None)
    assert L.lib.gscf_b200_engine_create(C.byref(cfg), C.byref(h)) == L.ERR_CUDA
    assert "no CPU fallback" in L.last_error()
    assert L.lib.gscf_b200_dgemm(0, 0, 1, 1, 1, 1.0, None, 1, 0, None, 1, 0, 0.0, None, 1, 0, 1, None) == L.ERR_CUDA


def test_host_side_scalars():
    from gscf_b200 import lib as L
    from oracle import compute_damping_coeff

    for s, c in [(1e-15, -1.0), (-1.0, 1e-15), (0.5, 1.0), (-1.0, 0.25), (-1.0, 1.0), (-0.01, 1.0), (-1.0, 0.51),
                 (-0.3, 0.4), (-2.0, 1.7)]:
        assert L.lib.gscf_b200_compute_damping_coeff(s, c, 0.05) == compute_damping_coeff(s, c, 0.05)
    p = L.Params()
    L.lib.gscf_b200_default_params(C.byref(p))
    assert (p.max_iter, p.n_eigenpairs, p.max_error_norm, p.diis_n_prev_steps, p.toda_n_prev_steps,
            p.toda_min_fock_prefactor) == (100, -1, 5e-7, 4, 2, 0.05)
    assert L.lib.gscf_b200_padded_ld(14) == 16 and L.lib.gscf_b200_padded_ld(1024) == 1024


def test_genmap_and_control_params():
    import gscf_b200 as G

    m = G.GenMap({"max_iter": 7, "eigensolver/method": "x", "n_prev_steps": 9})
    assert m.at("max_iter", 1) == 7 and m.at("nope", 3) == 3 and m.exists("max_iter")
    assert m.submap("eigensolver") == {"method": "x"}
    d = G.PulayDiisScf(m)
    assert d.max_iter == 7 and d.n_prev_steps == 4  # unknown key "n_prev_steps" ignored
    t = G.TruncatedOptDampScf({"toda_min_fock_prefactor": 0.1})
    assert t.min_fock_prefactor == 0.1 and t.n_prev_steps == 2 and t.max_error_norm == 5e-7
    assert G.PulayDiisScfKeys.n_prev_steps == "diis_n_prev_steps"
    assert G.TruncatedOptDampScfKeys.n_prev_steps == "toda_n_prev_steps"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (-pedantic) and a C host must link against the library
    (host-only entry points here: default parameters and the TODA line search, TruncatedOptDampScf.hh:370-394)."""
    import subprocess

    src = tmp_path / "c_host.c"
    src.write_text(
        '#include <stdio.h>\n#include "gscf_b200.h"\n'
        "int main(void) {\n"
        "  gscf_b200_params p;\n"
        "  gscf_b200_default_params(&p);\n"
        "  if (p.max_iter != 100 || p.diis_n_prev_steps != 4 || p.toda_n_prev_steps != 2) return 1;\n"
        "  if (p.max_error_norm != 5e-7 || p.toda_min_fock_prefactor != 0.05) return 2;\n"
        "  /* s < 0, c > 0: interior minimum -s / (2 c) = 0.25 */\n"
        "  if (gscf_b200_compute_damping_coeff(-1.0, 2.0, 0.05) != 0.25) return 3;\n"
        "  /* descent direction lost (s >= 0): full step like the reference */\n"
        "  if (gscf_b200_compute_damping_coeff(1.0, 2.0, 0.05) != 1.0) return 4;\n"
        '  printf("%s\\n", gscf_b200_last_error() ? "ok" : "null");\n'
        "  return 0;\n}\n")
    exe = tmp_path / "c_host"
    libdir = os.path.join(ROOT, "gscf_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-L", libdir, "-lgscf_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
