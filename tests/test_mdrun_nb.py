"""The cluster-pair kernel of libfepb200.so as a drop-in inside the reference's own mdrun, on the B200.

`GMX_EMULATE_GPU=1` makes the reference build its GPU-layout pair lists and evaluate them with nbnxn_kernel_gpu_ref
(src/gromacs/nbnxm/kerneldispatch.cpp:479); with GMX_FEPB200_NB=1 the hook of
integration/gromacs_shim/kerneldispatch_fepb200.patch hands the same lists, masked atoms and coordinates to
fepb200_nb_* instead.  Same binary, same run input, compared step by step like tests/test_nb_shim_cpu.py does on CPU
against the stand-in; the last test also routes the perturbed pairs through fepb200_* (GMX_FEPB200=1): then every
short-range non-bonded interaction of the run is computed by the library."""
import os

import pytest

import test_mdrun_dropin as T
from test_nb_shim_cpu import EMU, compare_nb_runs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(T.GMX), reason="integration/_gmx not built")]


def _wall_ms(workdir, name):
    """ms per call of a wall-cycle counter row of md.log, e.g. 'Force' or 'Nonbonded F'."""
    for line in open(os.path.join(workdir, "run.log")):
        if line.startswith(" " + name + " "):
            parts = line[len(name) + 1:].split()
            try:
                return 1e3 * float(parts[3]) / int(parts[2])
            except (IndexError, ValueError):
                return None
    return None


@pytest.mark.parametrize("system", ["coulandvdwtogether", "c1_methane", "c2_hexadecane", "transformAtoB", "vdwalone"])
def test_cluster_pairs_through_the_library_inside_mdrun(system, tmp_path):
    tpr = os.path.join(T.TPR, system + ".tpr")
    a, b = str(tmp_path / "ref"), str(tmp_path / "lib")
    ref = T._run(tpr, a, False, extra_env=EMU, mdrun_args=("-nstlist", "5"))
    via = T._run(tpr, b, False, extra_env=dict(EMU, GMX_FEPB200_NB="1"), mdrun_args=("-nstlist", "5"))
    assert "fepb200 nb shim:" in via[0] and "STAND-IN" not in via[0]
    compare_nb_runs(system, ref, via, (a, b))
    out = os.path.join(T.ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "mdrun_nb_dropin_timing.txt"), "a") as fh:
        shim = [ln for ln in via[0].splitlines() if ln.startswith("fepb200 nb shim:")]
        fh.write(f"{system}: md.log 'Nonbonded F' per call: reference kernel_gpu_ref {_wall_ms(a, 'Nonbonded F')} ms, "
                 f"through libfepb200 {_wall_ms(b, 'Nonbonded F')} ms ('Force': {_wall_ms(a, 'Force')} vs {_wall_ms(b, 'Force')} ms); {shim[-1] if shim else ''}\n")


def test_all_short_range_pairs_through_the_library(tmp_path):
    system = "c2_hexadecane"
    tpr = os.path.join(T.TPR, system + ".tpr")
    a, b = str(tmp_path / "ref"), str(tmp_path / "lib")
    ref = T._run(tpr, a, False, extra_env=EMU, mdrun_args=("-nstlist", "5"))
    via = T._run(tpr, b, True, extra_env=dict(EMU, GMX_FEPB200_NB="1"), mdrun_args=("-nstlist", "5"))
    assert "fepb200 nb shim:" in via[0] and "computed by fepb200" in via[0]
    compare_nb_runs(system, ref, via, (a, b))
