"""The reference-side shim (integration/gromacs_shim/fepb200_shim.h) checked without a GPU.

The reference's patched mdrun runs twice: perturbed pairs on its own CPU kernel, and through the shim
into tests/shim_standin/ -- a test-only stand-in for libfepb200.so that answers the nine entry points
the shim binds with the CPU oracle.  What is under test is the SHIM: which arrays it hands over and
when (atoms and pair list on search steps only), flag assembly, and where the results are added
(forces, energy-group terms, dvdl_lin / dvdl_nonlin, ForeignLambdaTerms).  The stand-in is not a
product path: libfepb200.so itself has no CPU route (tests/test_abi.py).

Needs integration/_gmx (integration/build_patched_gmx.sh; needs /root/reference)."""
import os
import re
import subprocess

import pytest

import test_mdrun_dropin as T

HERE = os.path.dirname(os.path.abspath(__file__))
STANDIN = os.path.join(HERE, "shim_standin", "libfepb200_standin.so")

pytestmark = pytest.mark.skipif(not os.path.exists(T.GMX), reason="integration/_gmx not built")


@pytest.fixture(scope="module")
def standin():
    src = os.path.join(HERE, "shim_standin", "fepb200_standin.c")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-D_POSIX_C_SOURCE=199309L", "-Wno-alloc-size-larger-than", "-Wno-stringop-overflow",
                           "-I", os.path.join(T.ROOT, "include"), "-o", STANDIN, src,
                           os.path.join(T.ROOT, "oracle", "fep_oracle.c"), "-lm"])
    return STANDIN


@pytest.mark.parametrize("system", ["coulandvdwtogether", "transformAtoB", "c1_methane", "c1_methane_ljpme", "c2_hexadecane",
                                    "c2_hexadecane_gapsys", "c2_hexadecane_rf", "coulandvdwintramol", "expanded", "relative",
                                    "relative-position-restraints",
                                    # slow growth: lambda moves every step (0.5 + 0.005 per step)
                                    "coulandvdwtogether_slowgrowth"])
def test_shim_hands_over_what_the_cpu_route_gets(system, standin, tmp_path):
    tpr = os.path.join(T.TPR, system + ".tpr")
    # mdrun would raise nstlist to 100 for these small systems: keep a pair search every 5 steps
    cpu = T._run(tpr, str(tmp_path / "cpu"), False, mdrun_args=("-nstlist", "5"))
    via = T._run(tpr, str(tmp_path / "shim"), True, lib=standin, pairs14=True, extra_env={"FEPB200_STANDIN_TRACE": "1"},
                 mdrun_args=("-nstlist", "5"))
    assert "CPU STAND-IN" in via[0]
    T.compare_runs(system, cpu, via)
    # cadence: nsteps steps with nstlist = 5 are nsteps + 1 force calls and nsteps / 5 + 1 pair searches
    # (20 steps: 21 calls, searches at steps 0, 5, ..., 20; the expanded-ensemble system runs 100 steps and
    # changes lambda on the way, which must reach the library without a new list)
    last = [ln for ln in via[0].splitlines() if ln.startswith("standin: compute")][-1]
    n = dict(zip(("compute", "set_list", "set_atoms", "set_params", "set_lambdas"), map(int, re.findall(r"\d+", last))))
    assert n["compute"] >= 21
    searches = (n["compute"] - 1) // 5 + 1
    assert n["set_list"] == n["set_atoms"] == searches, n
    assert n["set_params"] <= searches, n
    if system == "coulandvdwtogether_slowgrowth":
        assert n["set_lambdas"] == n["compute"], n  # a new lambda every step
    elif system != "expanded":
        assert n["set_lambdas"] <= searches, n
    assert "fepb200 shim:" in via[0]  # the shim's own timing summary at exit
    # the perturbed 1-4 pairs of the 50-atom solute go through fepb200_pairs14_* (hook in listed_forces/pairs.cpp;
    # compare_runs above includes LJ-14 and Coulomb-14): one hand-over per pair list a thread sees -- its chunk of the
    # force evaluation and, on the master thread, the list of the foreign-lambda evaluations
    p14 = [ln for ln in via[0].splitlines() if ln.startswith("fepb200 pairs14 shim:")]
    if system.startswith("c2_hexadecane"):
        # never more hand-overs than pair searches (mdrun sets the bonded threading up again on search steps)
        assert p14 and all(int(re.search(r"(\d+) pair-list uploads", ln).group(1)) <= searches for ln in p14), p14
        assert any("135 perturbed 1-4 pairs" in ln for ln in p14), p14
        # the foreign-lambda evaluations of a dH/dlambda step come from ONE library call (ForeignScope, hook in
        # listed_forces.cpp): n_lambda + 1 evaluations served per call
        m = [re.search(r"foreign lambda: (\d+) library calls served (\d+) evaluations", ln) for ln in p14]
        calls, served = max((int(x.group(1)), int(x.group(2))) for x in m if x)
        assert calls >= 5 and served % calls == 0 and served // calls >= 5, p14  # 21 / 9 / 41 lambda points per call
    elif system.startswith("coulandvdw") or system.startswith("c1_"):
        assert not p14  # no perturbed 1-4 pairs in these systems


def test_shim_with_domain_decomposition_one_context_per_rank(standin, tmp_path):
    """Two thread-MPI ranks: every rank has its own local + non-local FEP lists and local atom numbering,
    and the shim keeps one library context per rank (thread)."""
    tpr = os.path.join(T.TPR, "c2_hexadecane.tpr")
    args = ("-nstlist", "5", "-dd", "2", "1", "1")
    cpu = T._run(tpr, str(tmp_path / "cpu"), False, mdrun_args=args, ntmpi=2)
    via = T._run(tpr, str(tmp_path / "shim"), True, lib=standin, pairs14=True, mdrun_args=args, ntmpi=2)
    assert via[0].count("CPU STAND-IN") == 2  # one context per rank
    assert via[0].count("fepb200 shim:") == 2
    T.compare_runs("c2_hexadecane, 2 ranks", cpu, via)


@pytest.mark.parametrize("system", ["coulandvdwsequential_coul", "coulandvdwsequential_vdw", "coulandvdwtogether", "expanded",
                                    "relative", "relative-position-restraints", "transformAtoB", "vdwalone"])
def test_shim_route_reproduces_the_reference_golden_vectors(system, standin, tmp_path):
    """The run through the shim against the reference's OWN golden vectors of its mdrun free-energy test
    (tests/golden/mdrun_fe_refdata.json), at that test's tolerance.  With mdrun's default pair-list settings,
    as in the reference's test: with -nstlist 5 the reference's own CPU route moves away from its golden
    dV/dl by 0.57 kJ/mol on transformAtoB (the shim route moves with it)."""
    tpr = os.path.join(T.TPR, system + ".tpr")
    via = T._run(tpr, str(tmp_path / "shim"), True, lib=standin, pairs14=True)
    assert "CPU STAND-IN" in via[0]
    assert T.compare_with_reference_golden(system, via) >= 42
