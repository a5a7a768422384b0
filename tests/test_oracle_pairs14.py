"""Pins the CPU restatement of the perturbed 1-4 pair interactions (oracle/fep_oracle.c,
fep_oracle_pairs14) to the golden vectors of the reference's own test of do_pairs()
(tests/golden/pairs14_kat.json, from listed_forces/tests/refdata).  The reference evaluates the
interactions through spline tables and accepts 1e-7 (double); the restatement is analytic."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from pairs14_cases import kat_pairs14

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "pairs14_kat.json")) as fh:
    CASES = json.load(fh)["cases"]


@pytest.mark.parametrize("case", CASES, ids=[f"{c['pbc']}-{c['softcore']}-{c['lam']}" for c in CASES])
def test_pairs14_port_matches_reference_golden_vectors(case):
    out = oracle.run_pairs14(kat_pairs14(case))
    tol = 2e-7  # the reference's own double-precision tolerance is 1e-7 (table interpolation)
    f = np.array(case["forces"])
    assert np.max(np.abs(out["f"] - f)) <= tol * max(np.max(np.abs(f)), 1e-3)
    for got, want in ((out["Vc"][0], case["ECoul14"]), (out["Vv"][0], case["ELJ14"]), (out["dvdl"][0], case["dVdlCoul"]),
                      (out["dvdl"][1], case["dVdlVdw"])):
        assert abs(got - want) <= tol * max(abs(want), 1e-3), (got, want)
    assert np.max(np.abs(out["fshift"][22] - np.array(case["shift_force_central"]))) <= tol * max(np.max(np.abs(f)), 1e-3)
