"""Device-resident force hand-off (fepb200_add_forces_device) through the C-ABI.

This file sorts last on purpose: the entry point was written after round 1's GPU budget was spent, so
these tests have not run on a B200 yet; under `pytest -x` a surprise here must not hide the parity
tests that have."""
import numpy as np
import pytest

from fepb200 import params as P
from fepb200.synth import make_system, scaled_spec

pytestmark = pytest.mark.gpu

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA


@pytest.fixture(scope="module")
def ctx():
    from fepb200.lib import FepContext

    c = FepContext(0)
    yield c
    c.close()


def test_forces_added_into_a_device_resident_array(ctx):
    import torch

    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    ctx.set_problem(prob)
    want = ctx.compute(prob.x, prob.shiftvec, ALL)
    d_x = torch.from_numpy(np.ascontiguousarray(prob.x)).cuda()
    torch.cuda.synchronize()
    ctx.gather_x_device(d_x.data_ptr(), prob.shiftvec)
    ctx.launch(ALL)
    d_f = torch.zeros((prob.natoms, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.add_forces_device(d_f.data_ptr())
    ctx.add_forces_device(d_f.data_ptr())  # accumulates, like the reference kernel into its force buffer
    ctx.wait()
    assert np.array_equal(d_f.cpu().numpy(), 2.0 * want["f"])
    ctx.add_forces_device(d_f.data_ptr(), P.CLEAR_OUTPUTS)
    ctx.wait()
    assert np.array_equal(d_f.cpu().numpy(), want["f"])
    # the scalars still come from download(); without DO_FORCE it skips the force copy
    only_scalars = ctx.download(ALL & ~(P.DO_FORCE | P.DO_SHIFTFORCE))
    assert not only_scalars["f"].any()
    assert np.array_equal(only_scalars["Vc"], want["Vc"]) and np.array_equal(only_scalars["dvdl"], want["dvdl"])


def test_hand_off_refuses_results_that_live_in_host_memory(ctx):
    import torch

    from fepb200.lib import FepError

    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    ctx.set_problem(prob)
    ctx.compute(prob.x, prob.shiftvec, ALL)  # the epilogue wrote this step's results into pinned host memory
    d_f = torch.zeros((prob.natoms, 3), dtype=torch.float32, device="cuda")
    with pytest.raises(FepError):
        ctx.add_forces_device(d_f.data_ptr())
