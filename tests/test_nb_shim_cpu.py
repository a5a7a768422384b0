"""The reference-side binding of the cluster-pair kernel (integration/gromacs_shim/fepb200_nb_shim.h, hook in
src/gromacs/nbnxm/kerneldispatch.cpp) checked without a GPU, inside the reference's own mdrun.

`GMX_EMULATE_GPU=1` makes the reference build GPU-layout cluster pair lists (masked perturbed atoms, perturbed pairs moved
to the FEP list) and evaluate them with its reference kernel nbnxn_kernel_gpu_ref (kerneldispatch.cpp:479).  Each system
runs twice from one binary: with that kernel, and -- GMX_FEPB200_NB=1 -- through the shim into the test-only stand-in
(tests/shim_standin/fepb200_nb_standin.c, answered by oracle/nb_oracle.c).  Under test is the SHIM: what it hands over
(list structures as the reference holds them, masked types / charges, xyzq coordinates), when (atoms and list on search
steps only), and where the results go (forces, shift forces -> pressure, Coulomb-(SR), LJ-(SR)).

Needs integration/_gmx (integration/build_patched_gmx.sh; needs /root/reference)."""
import os
import re
import subprocess

import numpy as np
import pytest

import test_mdrun_dropin as T

HERE = os.path.dirname(os.path.abspath(__file__))
STANDIN = os.path.join(HERE, "shim_standin", "libfepb200_standin.so")

pytestmark = pytest.mark.skipif(not os.path.exists(T.GMX), reason="integration/_gmx not built")


@pytest.fixture(scope="module")
def standin():
    d = os.path.join(HERE, "shim_standin")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-D_DEFAULT_SOURCE",
                           "-Wno-alloc-size-larger-than", "-Wno-stringop-overflow", "-I", os.path.join(T.ROOT, "include"),
                           "-o", STANDIN, os.path.join(d, "fepb200_standin.c"), os.path.join(d, "fepb200_nb_standin.c"),
                           os.path.join(T.ROOT, "oracle", "fep_oracle.c"), os.path.join(T.ROOT, "oracle", "nb_oracle.c"), "-lm"])
    return STANDIN


def pressure(workdir, gmx=None):
    gmx = gmx or T.GMX
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(os.path.dirname(os.path.dirname(gmx)), "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    e = subprocess.run([gmx, "-quiet", "energy", "-f", "run.edr", "-o", "pres.xvg"], cwd=workdir, env=env,
                       input="Pressure\nPres-XX\nPres-YY\nPres-ZZ\n\n", capture_output=True, text=True, timeout=120)
    assert e.returncode == 0, e.stderr[-1500:]
    return T._xvg(os.path.join(workdir, "pres.xvg"))


def compare_nb_runs(system, ref, via, workdirs, rtol=1e-4):
    """Per-step energies at the tolerance of the reference's own mdrun free-energy test (relative 1e-4 with an absolute floor
    of 50 * 1e-4 kJ/mol, src/programs/mdrun/tests/freeenergy.cpp:115-117), pressure (shift forces -> virial) at 1e-3 of its
    fluctuation scale, dH/dlambda output unchanged (the perturbed pairs are on the same route in both runs)."""
    _, terms_a, e_a, dh_a = ref
    _, terms_b, e_b, dh_b = via
    assert terms_a == terms_b and e_a.shape == e_b.shape and e_a.shape[0] >= 20
    for col, name in enumerate(terms_a, start=1):
        scale = max(np.max(np.abs(e_a[:, col])), 50.0)
        assert np.max(np.abs(e_b[:, col] - e_a[:, col])) <= rtol * scale, (system, name, np.max(np.abs(e_b[:, col] - e_a[:, col])), scale)
    for col in range(1, dh_a.shape[1]):
        scale = max(np.max(np.abs(dh_a[:, col])), 50.0)
        assert np.max(np.abs(dh_b[:, col] - dh_a[:, col])) <= rtol * scale, (system, "dh", col)
    p_a, p_b = pressure(workdirs[0]), pressure(workdirs[1])
    assert p_a.shape == p_b.shape and p_a.shape[1] >= 2
    scale = max(np.max(np.abs(p_a[:, 1:])), 100.0)  # bar
    assert np.max(np.abs(p_b[:, 1:] - p_a[:, 1:])) <= 2e-3 * scale, (system, "pressure", np.max(np.abs(p_b[:, 1:] - p_a[:, 1:])), scale)


EMU = {"GMX_EMULATE_GPU": "1"}


@pytest.mark.parametrize("system", ["coulandvdwtogether", "c1_methane", "c2_hexadecane", "transformAtoB", "vdwalone"])
def test_shim_hands_over_what_the_reference_kernel_gets(system, standin, tmp_path):
    tpr = os.path.join(T.TPR, system + ".tpr")
    a, b = str(tmp_path / "ref"), str(tmp_path / "shim")
    ref = T._run(tpr, a, False, extra_env=EMU, mdrun_args=("-nstlist", "5"))
    via = T._run(tpr, b, False, lib=standin, extra_env=dict(EMU, GMX_FEPB200_NB="1", FEPB200_STANDIN_TRACE="1"),
                 mdrun_args=("-nstlist", "5"))
    assert "fepb200_nb CPU STAND-IN" in via[0] and "fepb200_nb" not in ref[0]
    compare_nb_runs(system, ref, via, (a, b))
    last = [ln for ln in via[0].splitlines() if ln.startswith("nb standin: compute")][-1]
    n = dict(zip(("compute", "set_pairlist", "set_atoms", "set_params"), map(int, re.findall(r"\d+", last)[:4])))
    assert n["compute"] >= 21
    searches = (n["compute"] - 1) // 5 + 1
    assert n["set_pairlist"] == n["set_atoms"] == n["set_params"] == searches, n
    assert "fepb200 nb shim:" in via[0]  # the shim's timing summary at exit
    # energies / shift forces are asked for on the steps the reference computes them (these systems: every step), the
    # charges come from the reference's own masked xyzq array
    flags = {int(m, 16) for m in re.findall(r"flags (0x[0-9a-f]+)", via[0])}
    assert all(f & 0x2 for f in flags) and any(f & 0x10 for f in flags) and all(f & (1 << 20) for f in flags)


def test_both_bindings_at_once(standin, tmp_path):
    """Perturbed pairs through fepb200_* and the cluster pairs through fepb200_nb_* in one run."""
    system = "coulandvdwtogether"
    tpr = os.path.join(T.TPR, system + ".tpr")
    a, b = str(tmp_path / "ref"), str(tmp_path / "shim")
    ref = T._run(tpr, a, False, extra_env=EMU, mdrun_args=("-nstlist", "5"))
    via = T._run(tpr, b, True, lib=standin, extra_env=dict(EMU, GMX_FEPB200_NB="1"), mdrun_args=("-nstlist", "5"))
    assert "fepb200_nb CPU STAND-IN" in via[0] and "computed by fepb200" in via[0]
    compare_nb_runs(system, ref, via, (a, b))
