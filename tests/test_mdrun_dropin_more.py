"""The mdrun drop-in test (tests/test_mdrun_dropin.py) on more systems and a domain-decomposed case: BASELINE.json's configurations as real GROMACS systems, the rest of
the reference's mdrun free-energy test systems, two thread-MPI ranks.  All pass through the shim on CPU
(tests/test_shim_cpu.py: same hand-over, the fp64 oracle behind the entry points)."""
import os

import pytest

import test_mdrun_dropin as T
from test_mdrun_dropin import GMX, TPR, _run, compare_runs

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(GMX), reason="integration/_gmx not built (integration/build_patched_gmx.sh)")
@pytest.mark.parametrize("system", T.MORE_SYSTEMS)
def test_mdrun_with_the_library_matches_mdrun_with_the_reference_kernel(system, tmp_path):
    T.run_both_routes_and_compare(system, tmp_path)


@pytest.mark.skipif(not os.path.exists(GMX), reason="integration/_gmx not built (integration/build_patched_gmx.sh)")
def test_mdrun_with_two_domain_decomposition_ranks(tmp_path):
    """Two thread-MPI ranks (domain decomposition 2x1x1) sharing the B200: each rank has its own local and
    non-local FEP lists and its own library context (the shim keeps one per rank thread); with
    GMX_FEPB200_DEVICES=N the ranks would be spread over N GPUs.  (Added after round 1's GPU budget was
    spent: verified on CPU through tests/test_shim_cpu.py, first GPU run is the round-end one.)"""
    tpr = os.path.join(TPR, "c2_hexadecane.tpr")
    args = ("-nstlist", "5", "-dd", "2", "1", "1")
    cpu = _run(tpr, str(tmp_path / "cpu"), False, mdrun_args=args, ntmpi=2)
    gpu = _run(tpr, str(tmp_path / "gpu"), True, mdrun_args=args, ntmpi=2)
    assert gpu[0].count("computed by fepb200") == 2
    compare_runs("c2_hexadecane, 2 ranks", cpu, gpu)
