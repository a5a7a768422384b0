"""The GPU-route shim (integration/gromacs_shim/fepb200_gpu_shim.h) exercised on the CPU every round.

tests/shim_standin/gpu_shim_driver.cpp plays the fork's GPU route -- the calls the hooks of
integration/gromacs_shim/nbnxm_gpu_fepb200.patch make, at the fork's cadence, for a rank with two localities --
with host arrays standing in for the fork's NBAtomDataGpu and the test-only stand-in library (the fp64 oracle
behind the entry points) in place of libfepb200.so.  Checked: what the shim hands over and when (constants,
atoms, lambdas, one list per locality; again on search steps only), flag assembly, and that the library is only
allowed to add into the buffers the fork clears and copies back on that kind of step.  The real library behind
the same header runs in tests/test_mdrun_gpu_route.py (GPU)."""
import os
import subprocess

import numpy as np
import pytest

from fepb200 import params as P
from fepb200.synth import make_system, scaled_spec
from test_host_cpp import _write_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "shim_standin")


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    out = tmp_path_factory.mktemp("gpu_shim")
    lib, drv = str(out / "libfepb200_standin.so"), str(out / "gpu_shim_driver")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-D_POSIX_C_SOURCE=199309L",
                           "-Wno-alloc-size-larger-than", "-Wno-stringop-overflow", "-I", os.path.join(ROOT, "include"),
                           "-o", lib, os.path.join(HERE, "fepb200_standin.c"), os.path.join(ROOT, "oracle", "fep_oracle.c"), "-lm"])
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(HERE, "stub_include"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "integration", "gromacs_shim"),
                           "-o", drv, os.path.join(HERE, "gpu_shim_driver.cpp"), "-ldl"])
    return lib, drv


def _moved(x):
    i = np.arange(x.size, dtype=np.uint64)
    d = np.float32(0.003) * ((i * np.uint64(2654435761)) % np.uint64(7)).astype(np.float32) - np.float32(0.009)
    return (x.reshape(-1) + d).reshape(x.shape).astype(np.float32)


@pytest.mark.parametrize("name,localities", [("C2", 2), ("C4g1", 2), ("C2", 1)])
def test_gpu_route_shim_cadence_flags_and_routing(name, localities, built, tmp_path):
    from oracle import oracle

    lib, drv = built
    spec = {"C2": scaled_spec("C2", 3.6, 1, 30, n_foreign=5), "C4g1": scaled_spec("C4", 4.2, 2, 25, n_foreign=6, n_energy_groups=1)}[name]
    prob = make_system(spec)
    assert prob.nenergrp_pairs == 1  # the fork's GPU route has one energy group
    _write_problem(prob, 1, tmp_path / "p.bin")
    env = dict(os.environ, GMX_FEPB200="1", GMX_FEPB200_LIB=lib, FEPB200_STANDIN_TRACE="1")
    r = subprocess.run([drv, str(tmp_path / "p.bin"), str(tmp_path / "r.bin"), str(localities)], capture_output=True, text=True,
                       env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    # one context per locality (none for a locality whose list is empty); per context: constants once, lambdas when they change (set-up and step 2), atoms and
    # list on the two search steps, three launches
    notes = [ln for ln in r.stderr.splitlines() if "GPU route, locality" in ln]
    assert len(notes) == localities and all("computed by fepb200" in ln for ln in notes)
    last = [ln for ln in r.stderr.splitlines() if ln.startswith("standin: launch")][-localities:]
    for ln in last:
        assert ln.split()[1:] == "launch 3 set_list 2 set_atoms 2 set_params 1 set_lambdas 2 set_stream 1".split(), ln

    n, l = prob.natoms, prob.n_foreign
    raw = np.fromfile(tmp_path / "r.bin", dtype=np.float32)
    per_step = 3 * n + 4 + 4 * (l + 1) + 135
    assert raw.size == 3 * per_step
    ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    ref0 = oracle.run_best(prob, ALL)
    import copy

    moved = copy.copy(prob)
    moved.x = _moved(prob.x)
    moved.lambda_ = np.array(prob.lambda_, dtype=np.float32).copy()  # slow growth: the per-step hook of do_force
    moved.lambda_[[P.LAMBDA_COUL, P.LAMBDA_VDW]] += np.float32(0.125)
    ref2 = oracle.run_best(moved, ALL & ~P.DO_FOREIGNLAMBDA)

    def unpack(k):
        s = raw[k * per_step:(k + 1) * per_step]
        o = 3 * n
        d = dict(f=s[:o].reshape(n, 3), eLJ=s[o], eElec=s[o + 1], dvdlLJ=s[o + 2], dvdlElec=s[o + 3])
        o += 4
        for key in ("eLJF", "eElF", "dLJF", "dElF"):
            d[key] = s[o:o + l + 1]
            o += l + 1
        d["fShift"] = s[o:o + 135].reshape(45, 3)
        return d

    def close(a, b, rel=2e-6):
        b = np.asarray(b, float)
        return np.all(np.abs(np.asarray(a, float) - b) <= rel * max(np.max(np.abs(b)), 1.0))

    fscale = np.max(np.abs(ref0["f"]))
    s0, s1, s2 = unpack(0), unpack(1), unpack(2)
    # step 0: everything, summed over the two localities
    assert close(s0["f"], ref0["f"]) and close(s0["fShift"], ref0["fshift"], 2e-6 * fscale)
    assert close(s0["eElec"], ref0["Vc"].sum()) and close(s0["eLJ"], ref0["Vv"].sum())
    assert close(s0["dvdlElec"], ref0["dvdl"][0]) and close(s0["dvdlLJ"], ref0["dvdl"][1])
    assert close(s0["eLJF"], ref0["foreign_energy"]) and not s0["eElF"].any()
    assert close(s0["dElF"], ref0["foreign_dvdl"][:, 0]) and close(s0["dLJF"], ref0["foreign_dvdl"][:, 1])
    # step 1: forces only; the sentinels in the buffers the fork neither clears nor reads on such a step are intact
    assert close(s1["f"], ref0["f"])
    assert s1["eLJ"] == 7.0 and s1["eElec"] == 7.0 and s1["dvdlLJ"] == 7.0 and s1["dvdlElec"] == 7.0
    assert np.all(s1["fShift"] == 7.0)
    assert not (s1["eLJF"].any() or s1["eElF"].any() or s1["dLJF"].any() or s1["dElF"].any())
    # step 2: new coordinates, new hand-over, new lambda; no foreign lambdas this step
    assert close(s2["f"], ref2["f"]) and close(s2["fShift"], ref2["fshift"], 2e-6 * fscale)
    assert close(s2["eElec"], ref2["Vc"].sum()) and close(s2["eLJ"], ref2["Vv"].sum())
    assert close(s2["dvdlElec"], ref2["dvdl"][0]) and close(s2["dvdlLJ"], ref2["dvdl"][1])
    assert not (s2["eLJF"].any() or s2["dLJF"].any() or s2["dElF"].any())
    assert not np.array_equal(s2["f"], s0["f"])


def test_gpu_route_shim_is_inert_without_the_switch(built, tmp_path):
    """Without GMX_FEPB200 every hook returns before touching anything (the fork's own kernels run): the driver's
    step() would abort on the missing hand-over if the setters had not been no-ops -- they are, so it aborts."""
    lib, drv = built
    prob = make_system(scaled_spec("C2", 3.0, 1, 10, n_foreign=2))
    _write_problem(prob, 1, tmp_path / "p.bin")
    env = {k: v for k, v in os.environ.items() if k != "GMX_FEPB200"}
    env["GMX_FEPB200_LIB"] = lib
    r = subprocess.run([drv, str(tmp_path / "p.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True, env=env)
    assert r.returncode == 3 and "never handed over" in r.stderr
