"""The perturbed 1-4 pair interactions inside the reference's mdrun through fepb200_pairs14_* (SURVEY 8f-4 as a
drop-in): integration/gromacs_shim/fepb200_pairs14_shim.h, included by the reference's listed_forces/pairs.cpp
through pairs_fepb200.patch.  With GMX_FEPB200 set, do_pairs_general() hands the pairs that take its free-energy
branch to the library and skips them; the run must reproduce the reference route, LJ-14 and Coulomb-14 included, at
the tolerance of the reference's own mdrun free-energy test.  Systems: the 50-atom solute of BASELINE configs[1]
(135 perturbed 1-4 pairs) with Beutler / Gapsys soft-core and with reaction-field + sc-coul + two energy groups.

The hook passes on CPU every round (tests/test_shim_cpu.py, the fp64 oracle behind the entry points).  The other mdrun drop-in
tests keep this hook switched off (GMX_FEPB200_NO_PAIRS14) so that they test what they tested before."""
import os

import pytest

import test_mdrun_dropin as T

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(T.GMX), reason="integration/_gmx not built (integration/build_patched_gmx.sh)")
@pytest.mark.parametrize("system", ["c2_hexadecane", "c2_hexadecane_gapsys", "c2_hexadecane_rf"])
def test_mdrun_with_perturbed_14_pairs_through_the_library(system, tmp_path):
    T.run_both_routes_and_compare(system, tmp_path, pairs14=True)
