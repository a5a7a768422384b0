"""The reference fork built WITH its CUDA back end for sm_100 (integration/build_patched_gmx_cuda.sh ->
integration/_gmx_cuda) and the same 13-line hook: the deployment the fork is made for -- non-bonded
cluster pairs on the GPU (`mdrun -nb gpu`) -- with the perturbed pairs

  (a) on the reference's CPU kernel                       mdrun -nb gpu -fep cpu
  (b) through libfepb200.so on the same B200              mdrun -nb gpu -fep cpu, GMX_FEPB200=1
  (c) on the fork's own CUDA FEP kernels                  mdrun -nb gpu -fep gpu

(b) must reproduce (a) at the tolerance of the reference's own mdrun free-energy test; (c) is run for the
record: the fork's "Nobonded FEP kernel" GPU time is noted beside our shim's wall time per step
(gpurun_out/mdrun_gpu_build_timing.txt).  (c) failing or deviating is the fork's business and does not
fail this test."""
import os
import re

import pytest

import test_mdrun_dropin as T

pytestmark = pytest.mark.gpu

GMX_CUDA = os.path.join(T.ROOT, "integration", "_gmx_cuda", "bin", "gmx")


def _fork_fep_gpu_ms_per_call(workdir):
    """'Nobonded FEP kernel' row of the GPU timing table of md.log (timing/wallcycle.cpp:1033-1037): count, total ms."""
    for line in open(os.path.join(workdir, "run.log")):
        if "FEP kernel" in line:
            nums = re.findall(r"[-+]?\d*\.\d+|\d+", line)
            if len(nums) >= 3:
                return line.strip()
    return None


@pytest.mark.skipif(not os.path.exists(GMX_CUDA), reason="integration/_gmx_cuda not built (integration/build_patched_gmx_cuda.sh)")
@pytest.mark.parametrize("system", ["coulandvdwtogether", "c2_hexadecane"])
def test_library_beside_the_forks_gpu_nonbonded_kernels(system, tmp_path):
    tpr = os.path.join(T.TPR, system + ".tpr")
    try:
        cpu = T._run(tpr, str(tmp_path / "a"), False, gmx=GMX_CUDA, nb="gpu", fep="cpu")
    except AssertionError as exc:  # the fork's GPU build itself does not run on this box: nothing of ours was involved yet
        pytest.skip("the fork's CUDA build does not run here: " + str(exc)[-400:])
    ours = T._run(tpr, str(tmp_path / "b"), True, gmx=GMX_CUDA, nb="gpu", fep="cpu")
    T.compare_runs(system, cpu, ours)
    T.compare_with_reference_golden(system, ours)
    note = [f"{system}: -nb gpu; perturbed pairs through libfepb200: "
            + next((ln for ln in ours[0].splitlines() if ln.startswith("fepb200 shim:")), "no shim summary")]
    a, b = T._nb_fep_ms_per_call(str(tmp_path / "a")), T._nb_fep_ms_per_call(str(tmp_path / "b"))
    note.append(f"{system}: NB FEP counter per call: reference CPU kernel (2 threads) {a} ms, libfepb200 {b} ms")
    try:  # (c): the fork's own GPU FEP route, for the record
        fork = T._run(tpr, str(tmp_path / "c"), False, gmx=GMX_CUDA, nb="gpu", fep="gpu")
        import numpy as np

        dev = {name: float(np.max(np.abs(fork[2][:, 1 + i] - cpu[2][:, 1 + i]))) for i, name in enumerate(cpu[1])}
        note.append(f"{system}: fork -fep gpu: max |dE| vs its CPU route {dev}; GPU timing row: "
                    f"{_fork_fep_gpu_ms_per_call(str(tmp_path / 'c'))}")
    except BaseException as exc:  # noqa: BLE001
        note.append(f"{system}: fork -fep gpu did not complete: {type(exc).__name__}: {str(exc)[:300]}")
    print("\n".join(note))
    try:
        out = os.path.join(T.ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "mdrun_gpu_build_timing.txt"), "a") as fh:
            fh.write("\n".join(note) + "\n")
    except OSError:
        pass
