"""Perturbed 1-4 pair interactions on the GPU (fepb200_pairs14_*) against the
reference's golden vectors (float tolerance of the reference's own test: 1e-5) and against the
double-precision oracle on random perturbed 1-4 lists."""
import json
import os

import numpy as np
import pytest

from fepb200 import params as P
from fepb200.pairs14 import PBC_NONE, PBC_XY, PBC_XYZ, Pairs14Context
from pairs14_cases import kat_pairs14, random_pairs14

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "pairs14_kat.json")) as fh:
    CASES = json.load(fh)["cases"]
FLAGS = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL


@pytest.fixture(scope="module")
def ctx():
    c = Pairs14Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", CASES, ids=[f"{c['pbc']}-{c['softcore']}-{c['lam']}" for c in CASES])
def test_reference_golden_vectors(ctx, case):
    prob = kat_pairs14(case, np.float32)
    ctx.set_problem(prob)
    out = ctx.compute(prob, FLAGS)
    tol = 1e-5  # pairs.cpp test: ListInput(1e-5, 1e-7)
    f = np.array(case["forces"])
    assert np.max(np.abs(out["f"] - f)) <= tol * max(np.max(np.abs(f)), 1e-3)
    for got, want in ((out["Vc"][0], case["ECoul14"]), (out["Vv"][0], case["ELJ14"]), (out["dvdl"][0], case["dVdlCoul"]),
                      (out["dvdl"][1], case["dVdlVdw"])):
        assert abs(got - want) <= tol * max(abs(want), 1e-3), (got, want)
    assert np.max(np.abs(out["fshift"][22] - np.array(case["shift_force_central"]))) <= tol * max(np.max(np.abs(f)), 1e-3)


@pytest.mark.parametrize("softcore", ["beutler", "gapsys"])
@pytest.mark.parametrize("sc_power,sc_coul", [(1, True), (2, True), (1, False)])
@pytest.mark.parametrize("pbc", [PBC_NONE, PBC_XYZ, PBC_XY])
def test_random_lists_match_oracle(ctx, softcore, sc_power, sc_coul, pbc):
    from oracle import oracle

    prob = random_pairs14(7 + sc_power + 3 * pbc, softcore, sc_power=sc_power, sc_coul=sc_coul, pbc_type=pbc)
    ctx.set_problem(prob)
    out = ctx.compute(prob, FLAGS)
    ref = oracle.run_pairs14(prob)
    rms = np.sqrt(np.mean((out["f"] - ref["f"]) ** 2) / np.mean(ref["f"] ** 2))
    assert rms <= 1e-5
    assert np.max(np.abs(out["fshift"] - ref["fshift"])) <= 1e-5 * np.sum(np.abs(ref["f"]), axis=0).max()
    for k in ("Vc", "Vv", "dvdl"):
        scale = np.maximum(np.abs(ref[k]), 1e-2 * np.max(np.abs(ref[k])))
        assert np.all(np.abs(out[k] - ref[k]) <= 1e-4 * scale), k


@pytest.mark.parametrize("softcore", ["beutler", "gapsys"])
@pytest.mark.parametrize("pbc", [PBC_NONE, PBC_XYZ])
def test_all_foreign_lambda_points_in_one_call(ctx, softcore, pbc):
    """fepb200_pairs14_compute_foreign: what the reference gets from one call of the pair code per lambda point
    (calc_listed_lambda, listed_forces.cpp:554-640) from ONE evaluation.  Energies against the oracle at every point;
    dV/dlambda against the energy-only evaluation of the same point (the semantics of the foreign passes of the
    non-bonded kernel: without the radius-derivative term of the Beutler soft-core, which is built from forces) and,
    for Gapsys -- whose dV/dlambda has no such term -- against the oracle as well."""
    import copy

    from oracle import oracle

    prob = random_pairs14(31 + pbc, softcore, sc_coul=True, pbc_type=pbc)
    ctx.set_problem(prob)
    lc = np.linspace(0.0, 1.0, 9).astype(np.float32)
    lv = (lc**2).astype(np.float32)  # separate coul / vdw paths
    e, d = ctx.compute_foreign(prob, lc, lv)
    for i in range(len(lc)):
        q = copy.copy(prob)
        q.lambda_ = prob.lambda_.copy()
        q.lambda_[P.LAMBDA_COUL], q.lambda_[P.LAMBDA_VDW] = lc[i], lv[i]
        ref = oracle.run_pairs14(q)
        want = float(np.sum(ref["Vc"]) + np.sum(ref["Vv"]))
        assert abs(e[i] - want) <= 1e-4 * max(abs(want), 1e-2 * np.max(np.abs(e))), (i, e[i], want)
        only_e = ctx.compute(q, P.DO_POTENTIAL)
        assert np.allclose(d[i], only_e["dvdl"], rtol=1e-5, atol=1e-5 * np.max(np.abs(d))), (i, d[i], only_e["dvdl"])
        if softcore == "gapsys":
            assert np.allclose(d[i], ref["dvdl"], rtol=1e-4, atol=1e-4 * np.max(np.abs(d))), (i, d[i], ref["dvdl"])
    # the plain evaluation still works afterwards, with its own lambda
    out = ctx.compute(prob, FLAGS)
    ref = oracle.run_pairs14(prob)
    assert np.sqrt(np.mean((out["f"] - ref["f"]) ** 2) / np.mean(ref["f"] ** 2)) <= 1e-5


def test_outputs_accumulate_and_flag_subsets(ctx):
    from oracle import oracle

    prob = random_pairs14(99, "beutler")
    ctx.set_problem(prob)
    one = ctx.compute(prob, FLAGS)
    two = ctx.compute(prob, FLAGS, out={k: v.copy() for k, v in one.items()})
    for k in one:
        assert np.allclose(two[k], 2 * one[k], rtol=1e-5, atol=1e-5 * np.max(np.abs(one[k]))), k
    only_e = ctx.compute(prob, P.DO_POTENTIAL)
    assert not np.any(only_e["f"]) and not np.any(only_e["fshift"])
    ref = oracle.run_pairs14(prob)
    assert np.allclose(only_e["Vc"], ref["Vc"], rtol=1e-4, atol=1e-4 * np.max(np.abs(ref["Vc"])))
    # lambda can change between steps without a new list
    prob.lambda_[:] = 0.9
    ctx_out = ctx.compute(prob, FLAGS)
    ref2 = oracle.run_pairs14(prob)
    assert np.sqrt(np.mean((ctx_out["f"] - ref2["f"]) ** 2) / np.mean(ref2["f"] ** 2)) <= 1e-5


def test_empty_list_and_errors(ctx):
    from fepb200.lib import FepError

    prob = random_pairs14(5, "beutler", npairs=4)
    prob.iatoms = prob.iatoms[:0]
    prob.gid = prob.gid[:0]
    ctx.set_problem(prob)
    out = ctx.compute(prob, FLAGS)
    assert not any(np.any(v) for v in out.values())
    bad = random_pairs14(5, "beutler", npairs=4)
    bad.iatoms[0, 1] = bad.natoms + 3
    with pytest.raises(FepError):
        ctx.set_problem(bad)
