#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference checkout.

Run ONCE in the build container (``/root/reference`` does not exist on the GPU
box):   python tests/golden/make_golden.py [/root/reference]

Produces, next to this script,
  sturmian14.npz             T, S, V0 (14x14) and ERI (196x196) base tables parsed from
                             src/gscfmock/Integrals/IntegralsSturmian14.cc:24,40,56,72
  functionality_test.json    guess, golden eigenvalues / eigenvectors / energies and
                             iteration bounds of tests/FunctionalityTest.cc:41-171
The tables are numeric test data (not code); they are parsed, never re-typed.
"""
import json
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def parse_matrix_tables(path):
    src = open(path).read()
    out = {}
    for name in ["m_t_bb_base", "m_s_bb", "m_v0_bb_base", "m_i_bbbb_base"]:
        m = re.search(r"IntegralsSturmian14::" + name + r"\s*\{", src)
        assert m, name
        i = m.end()
        depth, j = 1, i
        while depth:
            ch = src[j]
            depth += ch == "{"
            depth -= ch == "}"
            j += 1
        body = src[i : j - 1]
        rows = re.findall(r"\{([^{}]*)\}", body)
        mat = np.array([[float(x) for x in re.findall(NUM, r)] for r in rows])
        out[name] = mat
    return out


def parse_functionality_test(path):
    src = open(path).read()

    def block(after, opener="{", closer="}"):
        i = src.index(after) + len(after)
        i = src.index(opener, i) + 1
        depth, j = 1, i
        while depth:
            depth += src[j] == opener
            depth -= src[j] == closer
            j += 1
        return src[i : j - 1]

    evals = [float(x) for x in re.findall(NUM, block("eval_expected"))]
    evec_rows = re.findall(r"\{([^{}]*)\}", block("evec_expected"))
    evecs = [[float(x) for x in re.findall(NUM, r)] for r in evec_rows]

    def scalar(name):
        m = re.search(name + r"\s*=\s*(" + NUM + r")\s*;", src)
        return float(m.group(1))

    gold = dict(
        source="tests/FunctionalityTest.cc",
        z_charge=scalar("z_charge"),
        k_exp=scalar("k_exp"),
        n_alpha=int(scalar("n_alpha")),
        n_beta=int(scalar("n_beta")),
        eval_expected=evals,
        # evec_expected is a lazyten MultiVector initialiser list: row i lists component i of
        # every eigenvector, i.e. the table is the (n_elem x n_vectors) matrix, vector k = column k
        # (checked: column 0/1 reproduce the converged occupied orbitals, rows do not).
        evec_expected_rows=evecs,
        exp_energy_1e_terms=scalar("exp_energy_1e_terms"),
        exp_energy_coulomb=scalar("exp_energy_coulomb"),
        exp_energy_exchange=scalar("exp_energy_exchange"),
        exp_energy_total=scalar("exp_energy_total"),
        basetol=scalar("basetol"),
        n_iter_bound=dict(plain=15, diis=8, toda=13),  # FunctionalityTest.cc:113,141,171
    )
    assert len(evals) == 14 and len(evecs) == 14 and all(len(v) == 14 for v in evecs)
    # sanity: bounds really are those in the file
    assert "res.n_iter() < 15" in src and "res.n_iter() < 8" in src and "res.n_iter() < 13" in src
    return gold


def main():
    tabs = parse_matrix_tables(os.path.join(REF, "src/gscfmock/Integrals/IntegralsSturmian14.cc"))
    assert tabs["m_t_bb_base"].shape == (14, 14)
    assert tabs["m_s_bb"].shape == (14, 14)
    assert tabs["m_v0_bb_base"].shape == (14, 14)
    assert tabs["m_i_bbbb_base"].shape == (196, 196)
    np.savez_compressed(
        os.path.join(HERE, "sturmian14.npz"),
        t_bb_base=tabs["m_t_bb_base"],
        s_bb=tabs["m_s_bb"],
        v0_bb_base=tabs["m_v0_bb_base"],
        i_bbbb_base=tabs["m_i_bbbb_base"],
    )
    gold = parse_functionality_test(os.path.join(REF, "tests/FunctionalityTest.cc"))
    with open(os.path.join(HERE, "functionality_test.json"), "w") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote sturmian14.npz and functionality_test.json")


if __name__ == "__main__":
    main()
