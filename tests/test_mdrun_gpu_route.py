"""libfepb200.so INSIDE the fork's GPU route (SURVEY 8f-2 in full): `mdrun -nb gpu -fep gpu` of the fork built with
its CUDA back end (integration/build_patched_gmx_cuda.sh -> integration/_gmx_cuda) and the hooks of
integration/gromacs_shim/nbnxm_gpu_fepb200.patch.  With GMX_FEPB200 set, the launches of the fork's FEP kernels
(k_calc_nb_fep, k_calc_nb_fep_foreign; nbnxm/cuda/nbnxm_cuda.cu:755-851) are replaced by

    fepb200_gather_xq_device(adat->xq) -> fepb200_launch -> fepb200_add_forces_device(adat->f)
        -> fepb200_export_scalars_device(adat->eLJ, eElec, dvdl*, e*Foreign, dvdl*Foreign, fShift)

on the nbnxm stream of the locality: coordinates, forces and scalars stay on the device, the fork's copy-back and
reduction (nbnxm_gpu_data_mgmt.cpp:1117-1300, gpu_common.h:139-191) run unchanged.

  (a) mdrun -nb gpu -fep cpu                 perturbed pairs on the reference's CPU kernel
  (b) mdrun -nb gpu -fep gpu, GMX_FEPB200=1  perturbed pairs through libfepb200 inside the GPU route
  (c) mdrun -nb gpu -fep gpu                 the fork's own FEP kernels, for the record

(b) must reproduce (a) at the tolerance of the reference's own mdrun free-energy test.  The fork sums its GPU
energies in float together with the non-perturbed ones, so (b) is also compared with (c)'s deviation from (a) in
the note written to gpurun_out/mdrun_gpu_route_timing.txt."""
import os

import numpy as np
import pytest

import test_mdrun_dropin as T

pytestmark = pytest.mark.gpu

GMX_CUDA = os.path.join(T.ROOT, "integration", "_gmx_cuda", "bin", "gmx")
# the "GPU timings" table of md.log (timing/wallcycle.cpp:1009-1060) is only written with GPU timing on
TIMING = {"GMX_ENABLE_GPU_TIMING": "1"}


def _gpu_fep_row(workdir):
    for line in open(os.path.join(workdir, "run.log")):
        if "FEP kernel" in line:
            return " ".join(line.split())
    return None


def _max_dev(run, ref):
    return {name: float(np.max(np.abs(run[2][:, 1 + i] - ref[2][:, 1 + i]))) for i, name in enumerate(ref[1])}


@pytest.mark.skipif(not os.path.exists(GMX_CUDA), reason="integration/_gmx_cuda not built (integration/build_patched_gmx_cuda.sh)")
# the slow-growth system: lambda moves every step; the fork's GPU route uploads lambda once at set-up,
# so (c) is expected to drift away from (a) there, while (b) gets the current lambda through the hook in do_force
@pytest.mark.parametrize("system", ["coulandvdwtogether", "c2_hexadecane", "coulandvdwtogether_slowgrowth"])
def test_library_inside_the_forks_gpu_route(system, tmp_path):
    tpr = os.path.join(T.TPR, system + ".tpr")
    try:
        cpu = T._run(tpr, str(tmp_path / "a"), False, gmx=GMX_CUDA, nb="gpu", fep="cpu")
    except AssertionError as exc:  # the fork's GPU build itself does not run on this box: nothing of ours was involved yet
        pytest.skip("the fork's CUDA build does not run here: " + str(exc)[-400:])
    ours = T._run(tpr, str(tmp_path / "b"), True, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=TIMING)
    assert "GPU route" in ours[0], "the hook in gpu_launch_kernel was not reached"
    note = [f"{system}: -nb gpu -fep gpu through libfepb200; max |dE| vs the reference CPU FEP kernel {_max_dev(ours, cpu)}; "
            f"GPU timing row (our kernels inside the fork's fep_k timer): {_gpu_fep_row(str(tmp_path / 'b'))}"]
    note += [ln for ln in ours[0].splitlines() if ln.startswith("fepb200 GPU route:")]
    try:  # (c): the fork's own kernels on the same route, for the record
        fork = T._run(tpr, str(tmp_path / "c"), False, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=TIMING)
        note.append(f"{system}: fork's own FEP kernels: max |dE| vs its CPU route {_max_dev(fork, cpu)}; "
                    f"GPU timing row: {_gpu_fep_row(str(tmp_path / 'c'))}")
    except BaseException as exc:  # noqa: BLE001
        note.append(f"{system}: fork -fep gpu did not complete: {type(exc).__name__}: {str(exc)[:300]}")
    print("\n".join(note))
    try:
        out = os.path.join(T.ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "mdrun_gpu_route_timing.txt"), "a") as fh:
            fh.write("\n".join(note) + "\n")
    except OSError:
        pass
    T.compare_runs(system, cpu, ours)
    T.compare_with_reference_golden(system, ours)


@pytest.mark.skipif(not os.path.exists(GMX_CUDA), reason="integration/_gmx_cuda not built (integration/build_patched_gmx_cuda.sh)")
def test_gpu_route_steady_state_timing_beside_the_forks_kernels(tmp_path):
    """The fork's own GPU timing table ("Nobonded FEP kernel", timing/wallcycle.cpp:1032-1037: its fep_k timer around the
    FEP launches of gpu_launch_kernel) for libfepb200 inside the route and for the fork's kernels, on c2_hexadecane
    (BASELINE configs[1] as a real system), 600 steps with the counters reset half way so that context creation and the
    first list hand-over are not in the row.  A measurement, not a parity test: it fails only if a run fails."""
    tpr = os.path.join(T.TPR, "c2_hexadecane.tpr")
    args = ("-nsteps", "600", "-resethway")
    try:
        ours = T._run(tpr, str(tmp_path / "b"), True, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=TIMING, mdrun_args=args)
    except AssertionError as exc:
        pytest.skip("the fork's CUDA build does not run here: " + str(exc)[-400:])
    note = ["c2_hexadecane, 600 steps, counters reset half way (steady state):",
            f"  libfepb200 inside the fork's fep_k timer: {_gpu_fep_row(str(tmp_path / 'b'))}"]
    note += ["  " + ln for ln in ours[0].splitlines() if ln.startswith("fepb200 GPU route:")]
    try:
        T._run(tpr, str(tmp_path / "c"), False, gmx=GMX_CUDA, nb="gpu", fep="gpu", extra_env=TIMING, mdrun_args=args)
        note.append(f"  the fork's own FEP kernels:               {_gpu_fep_row(str(tmp_path / 'c'))}")
    except BaseException as exc:  # noqa: BLE001
        note.append(f"  fork -fep gpu did not complete: {type(exc).__name__}: {str(exc)[:300]}")
    print("\n".join(note))
    try:
        with open(os.path.join(T.ROOT, "gpurun_out", "mdrun_gpu_route_timing.txt"), "a") as fh:
            fh.write("\n".join(note) + "\n")
    except OSError:
        pass
    assert _gpu_fep_row(str(tmp_path / "b")) is not None
