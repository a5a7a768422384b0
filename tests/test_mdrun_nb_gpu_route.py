"""The cluster-pair kernel of libfepb200.so INSIDE the fork's GPU route (`mdrun -nb gpu`): with GMX_FEPB200_NB set the hook of
integration/gromacs_shim/nbnxm_gpu_nb_fepb200.patch replaces the launch of the fork's nbnxn_kernel_*_cuda
(nbnxm/cuda/nbnxm_cuda.cu:738-750) by fepb200_nb_launch_device on the nbnxm stream: coordinates (adat->xq, charges masked in
.w), shift vectors, forces (adat->f), shift forces and energies (adat->eLJ / eElec) stay in the fork's device buffers.

  (a) mdrun -nb gpu -fep gpu                                 the fork's own kernels (cluster pairs + perturbed pairs)
  (b) the same with GMX_FEPB200_NB=1                          cluster pairs through libfepb200
  (c) the same with GMX_FEPB200_NB=1 and GMX_FEPB200=1        EVERY short-range non-bonded pair through libfepb200, device-
                                                              resident, both kernels adding into the fork's adat->f
(b) and (c) must reproduce (a) step by step at the tolerance of the reference's own mdrun free-energy test; (c) also the
reference's CPU FEP route.  The fork's GPU timing table ("Nonbonded F kernel" rows of md.log) is noted for (a) and (b)."""
import os

import pytest

import test_mdrun_dropin as T
from test_mdrun_gpu_route import GMX_CUDA, TIMING
from test_nb_shim_cpu import compare_nb_runs

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(GMX_CUDA), reason="integration/_gmx_cuda not built")]


def _gpu_rows(workdir):
    rows = []
    for line in open(os.path.join(workdir, "run.log")):
        if "Nonbonded F" in line or "FEP kernel" in line or "Pruning kernel" in line:
            rows.append(" ".join(line.split()))
    return rows


# c1 / c2: potential shift (made here); coulandvdwtogether, transformAtoB: the reference's own test systems, force switch
@pytest.mark.parametrize("system", ["c1_methane", "c2_hexadecane", "coulandvdwtogether", "transformAtoB"])
def test_cluster_pairs_through_the_library_inside_the_forks_gpu_route(system, tmp_path):
    tpr = os.path.join(T.TPR, system + ".tpr")
    a, b, c = str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "c")
    try:
        fork = T._run(tpr, a, False, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=TIMING)
    except AssertionError as exc:
        pytest.skip("the fork's CUDA build does not run here: " + str(exc)[-400:])
    ours = T._run(tpr, b, False, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=dict(TIMING, GMX_FEPB200_NB="1"))
    assert "non-perturbed cluster pairs (GPU route) are computed by libfepb200" in ours[0]
    assert "fepb200 nb GPU route:" in ours[0]
    compare_nb_runs(system, fork, ours, (a, b))
    both = T._run(tpr, c, True, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=dict(TIMING, GMX_FEPB200_NB="1"))
    assert "non-perturbed cluster pairs (GPU route)" in both[0] and "GPU route" in both[0] and "fepb200 GPU route:" in both[0]
    compare_nb_runs(system, fork, both, (a, c))
    note = [f"{system}: mdrun -nb gpu -fep gpu, GPU timing rows of md.log",
            "  the fork's own kernels:          " + " | ".join(_gpu_rows(a)),
            "  cluster pairs through libfepb200: " + " | ".join(_gpu_rows(b)),
            "  all pairs through libfepb200:     " + " | ".join(_gpu_rows(c))]
    note += ["  " + ln for ln in both[0].splitlines() if ln.startswith("fepb200 nb GPU route:")]
    print("\n".join(note))
    try:
        with open(os.path.join(T.ROOT, "gpurun_out", "mdrun_nb_gpu_route_timing.txt"), "a") as fh:
            fh.write("\n".join(note) + "\n")
    except OSError:
        pass


def test_uncovered_flavour_stays_with_the_forks_kernel(tmp_path):
    """LJ-PME is a flavour of the fork's CUDA kernels that the library's cluster kernel does not have: the hook says so once
    and the fork's kernel runs."""
    tpr = os.path.join(T.TPR, "c1_methane_ljpme.tpr")
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    try:
        fork = T._run(tpr, a, False, gmx=GMX_CUDA, nb="gpu", fep="cpu")
    except AssertionError as exc:
        pytest.skip("the fork's CUDA build does not run here: " + str(exc)[-400:])
    ours = T._run(tpr, b, False, gmx=GMX_CUDA, nb="gpu", fep="cpu", extra_env={"GMX_FEPB200_NB": "1"})
    assert "the cluster pairs stay on the fork's kernel" in ours[0]
    compare_nb_runs("c1_methane_ljpme", fork, ours, (a, b))


@pytest.mark.parametrize("system", ["c2_hexadecane", "c3_hexadecane"])
def test_steady_state_timing_beside_the_forks_kernel(system, tmp_path):
    """The fork's own GPU timing table for its cluster kernel and for ours inside the same route -- c2_hexadecane (24.5 k atoms)
    and c3_hexadecane (the same solute in a 10 nm box, 100 k atoms) -- 600 steps with the counters reset half way (context
    creation and the first list hand-over are not in the rows).  A measurement: it fails only if a run fails.  Both kernels work
    on the fork's dynamically pruned device list (fepb200_nb_use_device_list)."""
    tpr = os.path.join(T.TPR, system + ".tpr")
    if not os.path.exists(tpr):
        pytest.skip(system + ".tpr not made (integration/systems/make_systems.py " + system + ")")
    args = ("-nsteps", "600", "-resethway")
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    try:
        T._run(tpr, a, False, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=TIMING, mdrun_args=args)
    except AssertionError as exc:
        pytest.skip("the fork's CUDA build does not run here: " + str(exc)[-400:])
    ours = T._run(tpr, b, True, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=dict(TIMING, GMX_FEPB200_NB="1"), mdrun_args=args)
    note = [system + ", 600 steps, counters reset half way (steady state), GPU timing rows of md.log:",
            "  the fork's own kernels:       " + " | ".join(_gpu_rows(a)),
            "  all pairs through libfepb200: " + " | ".join(_gpu_rows(b))]
    note += ["  " + ln for ln in ours[0].splitlines() if ln.startswith(("fepb200 nb GPU route:", "fepb200 GPU route:"))]
    print("\n".join(note))
    try:
        with open(os.path.join(T.ROOT, "gpurun_out", "mdrun_nb_gpu_route_timing.txt"), "a") as fh:
            fh.write("\n".join(note) + "\n")
    except OSError:
        pass
    assert _gpu_rows(b)


def test_two_domain_decomposition_ranks(tmp_path):
    """Two thread-MPI ranks: every rank has a local and a non-local pair list, each with its own stream, and all four kernels of
    a rank (cluster pairs and perturbed pairs of both localities) add into ONE adat->f with atomic operations.  All pairs
    through libfepb200 against the fork's own kernels, same decomposition."""
    system = "c2_hexadecane"
    tpr = os.path.join(T.TPR, system + ".tpr")
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    try:
        fork = T._run(tpr, a, False, gmx=GMX_CUDA, nb="gpu", fep="gpu", ntmpi=2)
    except AssertionError as exc:
        pytest.skip("the fork's CUDA build does not run two ranks here: " + str(exc)[-400:])
    both = T._run(tpr, b, True, gmx=GMX_CUDA, nb="gpu", fep="gpu", ntmpi=2, extra_env={"GMX_FEPB200_NB": "1"})
    assert both[0].count("fepb200 nb GPU route: locality 1") >= 1, "the non-local lists did not go through the library"
    compare_nb_runs(system, fork, both, (a, b))
