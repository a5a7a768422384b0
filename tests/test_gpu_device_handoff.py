"""Device-resident force hand-off (fepb200_add_forces_device) through the C-ABI."""
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


def test_nbat_style_hand_off_xyzq_in_forces_out_in_a_permuted_index_space(ctx):
    """What the nbnxm GPU module holds: atoms in grid order, coordinates as float4
    {x,y,z,q}, forces as float3 in the same order.  The pair list is handed over in that index space
    (the fork remaps it the same way, nbnxm_gpu_data_mgmt.cpp:763-787); coordinates are gathered from
    the xyzq array and forces added into the device force array without touching the host.  Compared
    with the same problem in the original order through fepb200_compute(): the pairs and their order
    are the same, so the results are bit-identical up to the permutation."""
    import copy

    import torch

    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    ctx.set_problem(prob)
    want = ctx.compute(prob.x, prob.shiftvec, ALL)

    rng = np.random.default_rng(7)
    n = prob.natoms
    new_of_old = rng.permutation(n).astype(np.int32)  # "grid" index of every atom
    old_of_new = np.argsort(new_of_old)
    grid = copy.copy(prob)
    grid.nblist = copy.copy(prob.nblist)
    grid.nblist.iinr = new_of_old[prob.nblist.iinr]
    grid.nblist.jjnr = new_of_old[prob.nblist.jjnr]
    for name in ("qA", "qB", "typeA", "typeB"):
        setattr(grid, name, np.ascontiguousarray(getattr(prob, name)[old_of_new]))
    grid.x = np.ascontiguousarray(prob.x[old_of_new])
    ctx.set_problem(grid)

    xq = np.zeros((n, 4), np.float32)
    xq[:, :3] = grid.x
    xq[:, 3] = 123.0  # must be ignored
    d_xq = torch.from_numpy(xq).cuda()
    d_f = torch.zeros((n, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.gather_xq_device(d_xq.data_ptr(), prob.shiftvec)
    ctx.launch(ALL)
    ctx.add_forces_device(d_f.data_ptr())
    ctx.wait()
    got = ctx.download(ALL & ~P.DO_FORCE)
    f_grid = d_f.cpu().numpy()
    # the touched-atom numbering (and with it the order of the per-atom sums) follows the index
    # space, so forces agree to rounding, not bit for bit
    scale = np.sqrt(np.mean(want["f"] ** 2))
    assert np.sqrt(np.mean((f_grid[new_of_old] - want["f"]) ** 2)) <= 1e-6 * scale
    for k in ("Vc", "Vv", "dvdl", "foreign_energy"):
        assert np.allclose(got[k], want[k], rtol=1e-6, atol=1e-6 * np.max(np.abs(want[k]))), k


def test_scalars_exported_into_the_forks_device_buffers(ctx):
    """fepb200_export_scalars_device: energies, dV/dlambda, foreign terms and shift forces added into float
    device buffers laid out like the outputs of the fork's NBAtomDataGpu (what a hook in its GPU route needs)."""
    import torch

    prob = make_system(scaled_spec("C4", 4.2, 2, 25, n_foreign=6))  # 16 energy-group pairs: summed on export
    ctx.set_problem(prob)
    want = ctx.compute(prob.x, prob.shiftvec, ALL)
    ctx.upload_x(prob.x, prob.shiftvec)
    ctx.launch(ALL)
    n_l = prob.n_foreign + 1
    buf = {k: torch.full((n,), 1.0, dtype=torch.float32, device="cuda")  # pre-filled: the export adds
           for k, n in dict(eLJ=1, eElec=1, dvdlLJ=1, dvdlElec=1, eLJForeign=n_l, eElecForeign=n_l, dvdlLJForeign=n_l,
                            dvdlElecForeign=n_l, fShift=135).items()}
    torch.cuda.synchronize()
    ctx.export_scalars_device(ALL, **{k: v.data_ptr() for k, v in buf.items()})
    ctx.wait()
    got = {k: v.cpu().numpy().astype(np.float64) - 1.0 for k, v in buf.items()}

    def close(a, b):
        b = np.asarray(b, float)
        return np.all(np.abs(a - b) <= 2e-6 * max(np.max(np.abs(b)), 1.0) + 2e-7 * np.abs(b) + 1e-6)

    assert close(got["eElec"][0], want["Vc"].sum()) and close(got["eLJ"][0], want["Vv"].sum())
    assert close(got["dvdlElec"][0], want["dvdl"][0]) and close(got["dvdlLJ"][0], want["dvdl"][1])
    assert close(got["eLJForeign"], want["foreign_energy"]) and not got["eElecForeign"].any()
    assert close(got["dvdlElecForeign"], want["foreign_dvdl"][:, 0]) and close(got["dvdlLJForeign"], want["foreign_dvdl"][:, 1])
    assert close(got["fShift"], want["fshift"].reshape(-1))


def test_atomic_outputs_give_the_same_sums(ctx):
    """FEPB200_ATOMIC_OUTPUTS: the additions into the caller's device buffers are atomic (a rank of the fork
    with two localities has the kernels of the other locality adding into the same NBAtomDataGpu from their
    own stream).  Every element still has one writer of ours, so the values are those of the plain additions."""
    import torch

    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    ctx.set_problem(prob)
    ctx.upload_x(prob.x, prob.shiftvec)
    ctx.launch(ALL)
    n_l = prob.n_foreign + 1
    sizes = dict(eLJ=1, eElec=1, dvdlLJ=1, dvdlElec=1, eLJForeign=n_l, eElecForeign=n_l, dvdlLJForeign=n_l,
                 dvdlElecForeign=n_l, fShift=135)
    out = []
    for extra in (0, P.ATOMIC_OUTPUTS):
        d_f = torch.full((prob.natoms, 3), 0.5, dtype=torch.float32, device="cuda")
        buf = {k: torch.full((n,), 1.0, dtype=torch.float32, device="cuda") for k, n in sizes.items()}
        torch.cuda.synchronize()
        ctx.add_forces_device(d_f.data_ptr(), extra)
        ctx.export_scalars_device(ALL | extra, **{k: v.data_ptr() for k, v in buf.items()})
        ctx.wait()
        out.append((d_f.cpu().numpy(), {k: v.cpu().numpy() for k, v in buf.items()}))
    assert np.array_equal(out[0][0], out[1][0]) and (out[0][0] != 0.5).any()
    for k in sizes:
        assert np.array_equal(out[0][1][k], out[1][1][k]), k


def test_atomic_outputs_with_a_second_stream_really_adding_into_the_same_array():
    """The case FEPB200_ATOMIC_OUTPUTS exists for (fork: two localities, nbnxm_cuda_kernel_utils.cuh:765-805): two
    contexts with their own streams add the forces of the same problem into ONE device array at the same time,
    many times over, on top of a non-zero initial value.  Floating-point addition of the same two summands is
    order-independent, so after every round the array must hold initial + 2 k f bit for bit; a lost update
    (plain read-modify-write) or an overwrite shows as a mismatch."""
    import torch

    from fepb200.lib import FepContext

    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    a, b = FepContext(0), FepContext(0)
    try:
        for c in (a, b):
            c.set_problem(prob)
            c.upload_x(prob.x, prob.shiftvec)
            c.launch(ALL)
        want = a.download(ALL)["f"]
        assert np.array_equal(want, b.download(ALL)["f"])
        d_f = torch.full((prob.natoms, 3), 0.5, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        rounds = 50
        for _ in range(rounds):  # queued back to back on two streams: the kernels overlap
            a.add_forces_device(d_f.data_ptr(), P.ATOMIC_OUTPUTS)
            b.add_forces_device(d_f.data_ptr(), P.ATOMIC_OUTPUTS)
        a.wait()
        b.wait()
        expect = np.full((prob.natoms, 3), 0.5, np.float32)
        for _ in range(rounds):
            expect = (expect + want) + want
        got = d_f.cpu().numpy()
        touched = a.touched_atoms()
        # (0.5 + f) + f in either order of the two adders is the same float; compare exactly on touched atoms
        assert np.array_equal(got[touched], expect[touched])
        untouched = np.ones(prob.natoms, bool)
        untouched[touched] = False
        assert np.all(got[untouched] == 0.5)
    finally:
        a.close()
        b.close()


def test_launch_on_a_stream_of_the_caller_s_is_ordered_against_the_context_s_stream(ctx):
    """fepb200_launch(ctx, flags, stream) with a stream that is not the context's: the staging copy of upload_x (context
    stream) must have landed before the kernels start, and the consumers (add_forces_device, download: context stream) must
    wait for the kernels.  A long-running kernel in front of the staging copy makes a missing dependency visible."""
    import torch

    prob = make_system(scaled_spec("C2", 3.2, 1, 20, n_foreign=4))
    ctx.set_problem(prob)
    want = ctx.compute(prob.x, prob.shiftvec, ALL)
    other = torch.cuda.Stream()
    x2 = prob.x + np.float32(0.01)
    want2 = ctx.compute(x2, prob.shiftvec, ALL)
    assert not np.array_equal(want["f"], want2["f"])
    d_f = torch.zeros((prob.natoms, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for x, w in ((prob.x, want), (x2, want2), (prob.x, want)):
        ctx.upload_x(x, prob.shiftvec)
        ctx.launch(ALL, stream=other.cuda_stream)
        ctx.add_forces_device(d_f.data_ptr(), P.CLEAR_OUTPUTS)
        got = ctx.download(ALL & ~P.DO_FORCE)
        ctx.wait()
        assert np.array_equal(d_f.cpu().numpy(), w["f"])
        assert np.array_equal(got["Vc"], w["Vc"]) and np.array_equal(got["foreign_energy"], w["foreign_energy"])
