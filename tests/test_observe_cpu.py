"""Observability helpers: the Mathematica-format writer and the iteration recorder, no GPU needed."""
import io
import json

import numpy as np


def _load():
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gscf_b200", "observe.py")
    spec = importlib.util.spec_from_file_location("gscf_b200_observe", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_mathematica_writer_format():
    obs = _load()
    w = obs.MathematicaWriter(io.StringIO(), precision=6)
    w.write("pulayerror3", 1.5e-7)
    w.write("evals1", [-4.25, 0.5])
    w.write("fock1", np.array([[1.0, -2.0], [0.0, 1e10]]))
    w.write_comment("x")
    lines = w.getvalue().splitlines()
    assert lines[0] == "pulayerror3 = 1.500000*^-7;"
    assert lines[1] == "evals1 = {-4.250000,5.000000*^-1};"
    assert lines[2] == "fock1 = {{1.000000,-2.000000},{0.000000,1.000000*^10}};"
    assert lines[3] == "(* x *)"


def test_iteration_recorder_roundtrip(tmp_path):
    obs = _load()

    class S:  # the members of ScfState the recorder reads
        def __init__(self, it):
            self._it = it
            self.energy_1e, self.energy_2e = -14.0 - it, 3.0
            self.last_error_norm = 10.0 ** -it
            self.diis_coefficients = np.array([0.25, 0.75]) if it > 1 else np.zeros(0)
            self.damping_coeff = 1.0

        def n_iter(self):
            return self._it

        @property
        def energy_total(self):
            return self.energy_1e + self.energy_2e

    rec = obs.IterationRecorder()
    for it in (1, 2, 3):
        rec.record(S(it), orben=np.arange(3.0))
    data = json.loads(rec.to_json(str(tmp_path / "t.json")))
    assert [d["iter"] for d in data] == [1, 2, 3]
    assert data[1]["diis_coefficients"] == [0.25, 0.75] and "diis_coefficients" not in data[0]
    rec.to_npz(str(tmp_path / "t.npz"))
    z = np.load(str(tmp_path / "t.npz"))
    np.testing.assert_allclose(z["energy_total"], [-12.0, -13.0, -14.0])
    np.testing.assert_allclose(z["orben_2"], [0.0, 1.0, 2.0])


def test_eigensolver_submap_selects_the_inner_solver():
    """ScfBase.hh:104: the "eigensolver" sub-map is handed to the inner eigensolver; its "method" key picks among this
    library's dense solvers (lazyten's own values all mean the automatic choice), flat and nested spellings alike."""
    import gscf_b200 as G
    from gscf_b200 import lib as L

    assert G.PulayDiisScf({"eigensolver": {"method": "jacobi"}}).eig_method == L.EIG_JACOBI
    assert G.PlainScf({"eigensolver/method": "tridiag_dc"}).eig_method == L.EIG_TRIDIAG_DC
    s = G.TruncatedOptDampScf({"eigensolver": G.GenMap({"method": "arpack", "tolerance": 1e-9})})
    assert s.eig_method == L.EIG_AUTO and s.eigensolver_params["tolerance"] == 1e-9
    assert G.PlainScf({}).eig_method == L.EIG_AUTO
    m = G.GenMap()
    G.PulayDiisScf({"eigensolver": {"method": "jacobi"}}).get_control_params(m)
    assert m["eigensolver/method"] == "jacobi"
