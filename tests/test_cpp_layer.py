"""The C++ host layer (include/gscf/*.hh + shims) as a drop-in: the reference's FunctionalityTest
re-expressed in tests/cpp/functionality_test.cc, compiled here with g++ and run on the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(HERE, "cpp", "functionality_test.bin")


def _build():
    from gscf_b200 import lib as L

    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
           os.path.join(HERE, "cpp", "functionality_test.cc"), "-o", BIN, "-L" + libdir, "-lgscf_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return BIN


def test_cpp_layer_compiles_and_links():
    """CPU: the header layer with the reference's class names compiles against the C ABI."""
    assert os.path.exists(_build())


def test_cpp_host_layer_probe_cpu():
    """CPU: hook-override detection (lazy trampolines), OperatorLinearCombination for ordinary and block-diagonal
    operators, the DIIS system layout and the default residual error - tests/cpp/hooks_probe.cc, no engine created."""
    from gscf_b200 import lib as L

    libdir = os.path.dirname(L.LIB_PATH)
    exe = os.path.join(HERE, "cpp", "hooks_probe.bin")
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "cpp", "hooks_probe.cc"),
           "-o", exe, "-L" + libdir, "-lgscf_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PROBE OK" in r.stdout, r.stdout + r.stderr


def test_reference_header_names_exist():
    for name in ["ScfBase", "ScfStateBase", "PlainScf", "PulayDiisScf", "TruncatedOptDampScf", "ScfProblemMatrix_i",
                 "FocklikeMatrix_i", "ScfBaseKeys", "PlainScfKeys", "PulayDiisScfKeys", "TruncatedOptDampScfKeys",
                 "OperatorLinearCombination", "version"]:
        assert os.path.exists(os.path.join(ROOT, "include", "gscf", name + ".hh")), name


@pytest.mark.gpu
def test_cpp_functionality_on_gpu(tmp_path):
    exe = _build()
    d = np.load(os.path.join(HERE, "golden", "sturmian14.npz"))
    binf = tmp_path / "sturmian14.bin"
    with open(binf, "wb") as f:
        for key in ["t_bb_base", "s_bb", "v0_bb_base", "i_bbbb_base"]:
            f.write(np.ascontiguousarray(d[key], dtype=np.float64).tobytes())
    g = json.load(open(os.path.join(HERE, "golden", "functionality_test.json")))
    txt = tmp_path / "golden.txt"
    with open(txt, "w") as f:
        f.write(" ".join(repr(x) for x in g["eval_expected"]) + "\n")
        for row in g["evec_expected_rows"]:
            f.write(" ".join(repr(x) for x in row) + "\n")
        f.write(" ".join(repr(g[k]) for k in ["exp_energy_1e_terms", "exp_energy_coulomb", "exp_energy_exchange",
                                                "exp_energy_total"]) + "\n")
        # the unrestricted mock (n_alpha = 2, n_beta = 1) through the oracle: iteration counts, energy, orbital energies
        from oracle import plain_scf, pulay_diis_scf, truncated_opt_damp_scf
        from oracle.mock_fock import IntegralsSturmian14, MockFockProblem, functionality_test_guess

        prob = MockFockProblem(IntegralsSturmian14(4.0, 1.0), 2, 1)
        g0 = functionality_test_guess()
        F0, e1, e2 = prob.fock(np.stack([g0, g0]))
        rp, rd, rt = (fn(prob, F0, e1, e2) for fn in (plain_scf, pulay_diis_scf, truncated_opt_damp_scf))
        f.write(f"{rp.n_iter} {rd.n_iter} {rt.n_iter} {rd.energy_total!r}\n")
        f.write(" ".join(repr(float(x)) for x in rd.orben.ravel()) + "\n")
    r = subprocess.run([exe, str(binf), str(txt)], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL CHECKS PASSED" in r.stdout
