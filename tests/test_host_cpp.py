"""The C++ host layer (gromacs-fep-gpu_b200/host): FreeEnergyDispatchGpu, the mirror of the reference's
FreeEnergyDispatch (src/gromacs/nbnxm/freeenergydispatch.cpp:312-413), driven from a C++ program
through the C-ABI and compared with the CPU oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from fepb200 import params as P
from fepb200.params import CParams
from fepb200.synth import make_system, scaled_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gromacs-fep-gpu_b200", "lib")
DRIVER = os.path.join(LIB, "dispatch_driver")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "gromacs-fep-gpu_b200", "csrc")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "gromacs-fep-gpu_b200", "host")], stdout=subprocess.DEVNULL)


def _write_problem(prob, n_groups, path):
    nb = prob.nblist
    with open(path, "wb") as fh:
        hdr = np.array([prob.natoms, prob.ntype, nb.nri, nb.nrj, prob.nenergrp_pairs, prob.n_foreign, n_groups, 0], np.int32)
        fh.write(hdr.tobytes())
        fh.write(bytes(prob.params.to_c()))
        for a, dt in ((prob.lambda_, np.float32), (prob.all_lambda_coul, np.float32), (prob.all_lambda_vdw, np.float32),
                      (prob.nbfp, np.float32), (prob.nbfp_grid, np.float32), (prob.x, np.float32), (prob.qA, np.float32),
                      (prob.qB, np.float32), (prob.typeA, np.int32), (prob.typeB, np.int32), (prob.shiftvec, np.float32),
                      (nb.iinr, np.int32), (nb.gid, np.int32), (nb.shift, np.int32), (nb.jindex, np.int32),
                      (nb.jjnr, np.int32), (nb.excl_fep, np.int32)):
            fh.write(np.ascontiguousarray(a, dtype=dt).tobytes())


def test_host_library_builds_and_fails_loudly_without_gpu(tmp_path):
    _build()
    assert os.path.exists(os.path.join(LIB, "libfepb200_host.so")) and os.path.exists(DRIVER)
    assert ctypes.sizeof(CParams) == 96  # the struct the driver reads from the problem file
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    prob = make_system(scaled_spec("C2", 3.0, 1, 10, n_foreign=2))
    _write_problem(prob, 1, tmp_path / "p.bin")
    r = subprocess.run([DRIVER, str(tmp_path / "p.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "fepb200 error -3" in r.stderr  # no device, no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("name,groups", [("C2", 1), ("C4", 4), ("C3", 1)])
def test_cpp_dispatch_matches_oracle(tmp_path, name, groups):
    _check_cpp_dispatch(tmp_path, name, groups, None)


@pytest.mark.parametrize("name,groups", [("C2", 1), ("C4", 4), ("C3", 1)])
def test_cpp_dispatch_host_logic_on_cpu(tmp_path, name, groups):
    """The same driver program with the test-only stand-in (tests/shim_standin: the oracle behind the entry
    points the host layer calls) pre-loaded in place of libfepb200.so: what runs here is the C++ class itself --
    flag assembly, dvdl_lin / dvdl_nonlin routing, ForeignLambdaTerms, the accumulate semantics."""
    src = os.path.join(ROOT, "tests", "shim_standin", "fepb200_standin.c")
    lib = os.path.join(ROOT, "tests", "shim_standin", "libfepb200_standin.so")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-D_POSIX_C_SOURCE=199309L",
                           "-Wno-alloc-size-larger-than", "-Wno-stringop-overflow", "-I", os.path.join(ROOT, "include"),
                           "-o", lib, src, os.path.join(ROOT, "oracle", "fep_oracle.c"), "-lm"])
    _check_cpp_dispatch(tmp_path, name, groups, dict(os.environ, LD_PRELOAD=lib))


def _check_cpp_dispatch(tmp_path, name, groups, env):
    from oracle import oracle

    _build()
    spec = {"C2": scaled_spec("C2", 3.6, 1, 30, n_foreign=5), "C4": scaled_spec("C4", 4.2, 2, 25, n_foreign=6),
            "C3": scaled_spec("C3", 4.0, 2, 25, n_foreign=4)}[name]
    prob = make_system(spec)
    _write_problem(prob, groups, tmp_path / "p.bin")
    r = subprocess.run([DRIVER, str(tmp_path / "p.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    n, g, l = prob.natoms, prob.nenergrp_pairs, prob.n_foreign
    raw = open(tmp_path / "r.bin", "rb").read()
    off = 0

    def take(count, dt):
        nonlocal off
        a = np.frombuffer(raw, dtype=dt, count=count, offset=off)
        off += a.nbytes
        return a

    f = take(3 * n, np.float32).reshape(n, 3)
    fshift = take(135, np.float32).reshape(45, 3)
    vc, vv = take(g, np.float64), take(g, np.float64)
    dvdl_lin, dvdl_nonlin = take(7, np.float64), take(7, np.float64)
    fe, fd = take(l + 1, np.float64), take(l + 1, np.float64)
    f2 = take(3 * n, np.float32).reshape(n, 3)
    dvdl2 = take(7, np.float64)
    assert off == len(raw)

    flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
    ref = oracle.run_best(prob, flags)
    rms = np.sqrt(np.mean((f - ref["f"]) ** 2) / np.mean(ref["f"] ** 2))
    assert rms < 1e-5
    assert np.max(np.abs(fshift - ref["fshift"])) < 5e-4 * np.max(np.abs(ref["f"]))
    for got, want in ((vc, ref["Vc"]), (vv, ref["Vv"]), (fe, ref["foreign_energy"]),
                      (fd, ref["foreign_dvdl"].sum(axis=1))):
        scale = np.maximum(np.abs(want), 1e-2 * np.max(np.abs(want)))
        assert np.all(np.abs(got - want) <= 1e-4 * scale)
    # soft-core is active in all three systems: dV/dlambda goes to dvdl_nonlin, nothing to dvdl_lin
    assert not np.any(dvdl_lin)
    got = dvdl_nonlin[[P.LAMBDA_COUL, P.LAMBDA_VDW]]
    assert np.all(np.abs(got - ref["dvdl"]) <= 1e-4 * np.maximum(np.abs(ref["dvdl"]), 1e-2 * np.max(np.abs(ref["dvdl"]))))
    assert not np.any(np.delete(dvdl_nonlin, [P.LAMBDA_COUL, P.LAMBDA_VDW]))
    # the plain force step gives the same forces and dV/dlambda
    assert np.array_equal(f2, f)
    # (different launch configuration of the pass, hence a different summation order of the partials)
    assert np.allclose(dvdl2, dvdl_nonlin, rtol=2e-6, atol=0)
