"""Pins the CPU checkers (test infrastructure under oracle/) before anything is compared with them.

  1. oracle/fep_oracle.c (our C restatement) against the 72 golden vectors of the reference's own
     unit test of this path (tests/golden/nb_free_energy_kat.json, extracted from
     src/gromacs/gmxlib/nonbonded/tests/refdata/*.xml), at the reference's double-precision
     tolerance (tests/nb_free_energy.cpp:433-435: relative 1e-8 ... we use 2e-8 + tiny absolute).
  2. oracle/_ref (the reference kernel compiled in place) against the same vectors, which proves
     the harness feeds the kernel the way the reference test does.
  3. the restatement against oracle/_ref on seeded random problems covering what the 72 cases do
     not: k_rf != 0, sc-power 2, lambda_coul != lambda_vdw, many entries, energy groups, all
     shift vectors, clamps, force-only / energy-only passes and the foreign-lambda loop.
No GPU involved.
"""
import json
import os

import numpy as np
import pytest

from fepb200 import params as P
from fepb200.synth import random_problem, make_system, scaled_spec
from kat_cases import KAT_FLAGS, NUM_CASES, kat_problem
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "nb_free_energy_kat.json")) as fh:
    GOLDEN = {c["index"]: c for c in json.load(fh)["cases"]}

needs_ref = pytest.mark.skipif(not oracle.have_ref("dp"), reason="oracle/_ref not built / host lacks AVX2")


def _check_kat(out, gold, rtol):
    def close(a, b, what):
        a, b = np.asarray(a, float), np.asarray(b, float)
        scale = max(np.max(np.abs(b)), 1e-3)
        assert np.max(np.abs(a - b)) <= rtol * scale, f"{what}: {a} vs {b}"

    close(out["Vv"][0], gold["EVdw"], "EVdw")
    close(out["Vc"][0], gold["ECoul"], "ECoul")
    close(out["dvdl"][0], gold["dVdlCoul"], "dVdlCoul")
    close(out["dvdl"][1], gold["dVdlVdw"], "dVdlVdw")
    close(out["f"], gold["forces"], "forces")
    # the test's only shift index is 0 and it reports it as the "Central" shift force
    close(out["fshift"][0], gold["shift_force_central"], "shift force")


@pytest.mark.parametrize("index", range(NUM_CASES))
def test_port_matches_reference_golden_vectors(index):
    prob = kat_problem(index, np.float64)
    out = oracle.run_port(prob, KAT_FLAGS)
    _check_kat(out, GOLDEN[index], 2e-8)


@needs_ref
@pytest.mark.parametrize("index", range(NUM_CASES))
@pytest.mark.parametrize("use_simd", [True, False])
def test_ref_build_matches_reference_golden_vectors(index, use_simd):
    prob = kat_problem(index, np.float64)
    out = oracle.run_ref(prob, KAT_FLAGS, precision="dp", use_simd=use_simd)
    _check_kat(out, GOLDEN[index], 2e-8)


def _param_grid():
    grid = []
    for sc in ("beutler", "gapsys"):
        for coul, vdw, mod in (("pme", "cut", "potshift"), ("rf", "cut", "potshift"), ("cut", "cut", "potswitch"),
                               ("pme", "pme", "potshift"), ("pme", "cut", "forceswitch"), ("rf", "cut", "none")):
            for power in (1, 2):
                for sccoul in (False, True):
                    grid.append((sc, coul, vdw, mod, power, sccoul))
    return grid


def _compare(a, b, rtol, what=""):
    for key in ("f", "fshift", "Vc", "Vv", "dvdl", "foreign_energy", "foreign_dvdl"):
        x, y = np.asarray(a[key]), np.asarray(b[key])
        scale = max(np.max(np.abs(y)) if y.size else 0.0, 1e-6)
        err = np.max(np.abs(x - y)) if y.size else 0.0
        assert err <= rtol * scale, f"{what} {key}: max err {err:g} scale {scale:g}"


@needs_ref
@pytest.mark.parametrize("case,combo", list(enumerate(_param_grid())))
def test_port_matches_ref_build_on_random_problems(case, combo):
    sc, coul, vdw, mod, power, sccoul = combo
    prm = P.make_params(coulombtype=coul, vdwtype=vdw, vdw_modifier=mod, rvdw_switch=0.8 if "switch" in mod else 0.0,
                        softcore=sc, sc_alpha=0.5, sc_power=power, sc_coul=sccoul)
    # Well-conditioned problems (no overlapping atoms): the result does not depend on the order of
    # summation, so the restatement must agree with the scalar AND the SIMD flavour of the reference
    # for any split of the list over threads.
    prob = random_problem(1000 + 17 * case, prm, n_foreign=5, frac_overlap=0.0, dtype=np.float64)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    for use_simd in (False, True):
        ref = oracle.run_ref(prob, flags, precision="dp", use_simd=use_simd, nthreads=3)
        port = oracle.run_port(prob, flags, nthreads=2)
        _compare(port, ref, 5e-9, f"simd={use_simd}")


@needs_ref
@pytest.mark.parametrize("case,combo", list(enumerate(_param_grid()))[::3])
def test_port_matches_ref_build_with_overlapping_atoms(case, combo):
    """Atoms 1e-7 .. 2e-3 nm apart exercise the r^2 and r^-6 clamps (nb_free_energy.cpp:99,107).
    Terms of 1e25 then cancel inside dV/dlambda, so sums depend on the summation order even in
    double: compare with the SCALAR reference flavour on ONE thread, whose order of additions the
    restatement shares, where agreement stays at rounding level."""
    sc, coul, vdw, mod, power, sccoul = combo
    prm = P.make_params(coulombtype=coul, vdwtype=vdw, vdw_modifier=mod, rvdw_switch=0.8 if "switch" in mod else 0.0,
                        softcore=sc, sc_alpha=0.5, sc_power=power, sc_coul=sccoul)
    prob = random_problem(3000 + 17 * case, prm, n_foreign=3, frac_overlap=0.05, dtype=np.float64)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    ref = oracle.run_ref(prob, flags, precision="dp", use_simd=False, nthreads=1)
    port = oracle.run_port(prob, flags, nthreads=1)
    # the terms that cancel are as large as the LJ energies: that is the scale of the rounding error
    big = max(np.max(np.abs(ref["Vv"])), np.max(np.abs(ref["Vc"])))
    for key in ("f", "fshift", "Vc", "Vv", "foreign_energy"):
        scale = max(np.max(np.abs(ref[key])), 1e-6)
        assert np.max(np.abs(port[key] - ref[key])) <= 5e-9 * scale, key
    for key in ("dvdl", "foreign_dvdl"):
        assert np.max(np.abs(port[key] - ref[key])) <= 5e-9 * max(np.max(np.abs(ref[key])), 1e-9 * big), key


@needs_ref
@pytest.mark.parametrize("flags", [P.DO_FORCE, P.DO_FORCE | P.DO_SHIFTFORCE, P.DO_POTENTIAL,
                                   P.DO_FORCE | P.DO_POTENTIAL])
def test_port_matches_ref_build_flag_subsets(flags):
    prm = P.make_params(coulombtype="pme", softcore="beutler", sc_alpha=0.5, sc_coul=True)
    prob = random_problem(77, prm, dtype=np.float64, lambda_coul=0.5, lambda_vdw=0.5)
    ref = oracle.run_ref(prob, flags, precision="dp", use_simd=False)
    port = oracle.run_port(prob, flags)
    _compare(port, ref, 5e-9)


@needs_ref
def test_port_matches_ref_build_equal_lambdas_and_endpoints():
    for lam in (0.0, 1.0, 0.25):
        for sc in ("beutler", "gapsys"):
            prm = P.make_params(coulombtype="pme", softcore=sc, sc_alpha=0.5)
            prob = random_problem(5, prm, dtype=np.float64, lambda_coul=lam, lambda_vdw=lam, n_foreign=3)
            flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
            _compare(oracle.run_port(prob, flags), oracle.run_ref(prob, flags, precision="dp"), 5e-9, f"{sc} {lam}")


@needs_ref
def test_port_matches_ref_build_on_solvated_system():
    spec = scaled_spec("C2", 3.2, 1, 20, n_adversarial=4)
    prob = make_system(spec, dtype=np.float64)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    ref = oracle.run_ref(prob, flags, precision="dp", nthreads=4)
    port = oracle.run_port(prob, flags, nthreads=4)
    _compare(port, ref, 5e-9)


@needs_ref
def test_mixed_precision_reference_is_close_to_double():
    """The error budget of BASELINE.json: the reference's own float build vs its double build."""
    if not oracle.have_ref("sp"):
        pytest.skip("no single-precision reference library for this host")
    spec = scaled_spec("C2", 3.2, 1, 20, n_adversarial=0)
    prob = make_system(spec)
    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    dp = oracle.run_ref(prob, flags, precision="dp")
    sp = oracle.run_ref(prob, flags, precision="sp", nthreads=2)
    rms = np.sqrt(np.mean((sp["f"] - dp["f"]) ** 2)) / np.sqrt(np.mean(dp["f"] ** 2))
    assert rms < 1e-5
    assert np.allclose(sp["foreign_energy"], dp["foreign_energy"], rtol=1e-4, atol=1e-3)
