"""Parity of the non-perturbed cluster-pair kernel (include/fepb200_nb.h) with its pinned CPU checker
(oracle/nb_oracle.c, tests/test_oracle_nb.py), through the C-ABI, and the device-resident hand-off with the perturbed-pair
kernels: both add into ONE force buffer in grid order on the device.

Tolerances (fp32 pair maths against the fp64 oracle; the reference's own float build deviates by 1e-6 rel-RMS in the forces,
test_oracle_nb.py::test_float_build_of_the_reference_sets_the_error_budget):
  forces        rel-RMS <= 2e-6 of the oracle's (analytical Ewald on both sides), max deviation <= 2e-5 of the largest force
  shift forces  <= 2e-5 of the largest component
  energies      <= 1e-5 relative (sums of ~1e6 cancelling terms in fp32 lanes, fp64 across warps)
"""
import numpy as np
import pytest

from fepb200 import params as P
from fepb200 import synth_nb
from fepb200.synth import make_system, scaled_spec
from oracle import nb_oracle

pytestmark = pytest.mark.gpu

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL

CASES = {
    "ewald": dict(name="C2", box=4.2, n_blobs=1),
    "ewald_no_overlaps": dict(name="C2", box=4.2, n_blobs=1, n_adversarial=0),  # every sum and energy at its ordinary size
    "rf_no_overlaps": dict(name="C4", box=4.2, n_blobs=1, n_energy_groups=1, n_adversarial=0),
    "rf": dict(name="C4", box=4.2, n_blobs=1, n_energy_groups=1),
    "ewald_split_entries": dict(name="C1", box=3.6, n_blobs=1, split=3),
    "ewald_larger": dict(name="C3", box=6.0, n_blobs=2),
}


def _system(name, box, n_blobs, seed=5, split=0, **kw):
    pr = make_system(scaled_spec(name, box, n_blobs, **kw), seed=seed)
    return pr, synth_nb.build_cluster_system(pr, rlist=1.1, max_cj_groups_per_sci=split)


@pytest.fixture(scope="module")
def nb():
    from fepb200.nb import NbContext

    c = NbContext(0)
    yield c
    c.close()


def _check(got, want, what="", core_tol=3e-6):
    f, fw = np.asarray(got["f"], np.float64), want["f"]
    rms = np.sqrt(np.mean((f - fw) ** 2) / np.mean(fw**2))
    assert rms <= 2e-6, f"{what}: force rel-RMS {rms:.2e}"
    assert np.max(np.abs(f - fw)) <= 2e-5 * np.max(np.abs(fw)), what
    # The synthetic systems carry adversarial placements (fepb200.synth: a water molecule moved next to a soft-cored atom),
    # which can put two waters on top of each other; one such pair (1e12 kJ/mol/nm) dominates the sums above.  The atoms
    # with ordinary forces are therefore checked on their own (measured 1.0e-6 ... 1.5e-6, with and without the
    # placements: profiles/r02_nb_parity_report.txt)
    mag = np.linalg.norm(fw, axis=1)
    core = mag <= 50 * np.median(mag[mag > 0])
    rms_core = np.sqrt(np.mean((f[core] - fw[core]) ** 2) / np.mean(fw[core] ** 2))
    assert rms_core <= core_tol, f"{what}: force rel-RMS over the atoms with ordinary forces {rms_core:.2e}"
    return rms


@pytest.mark.parametrize("case", sorted(CASES))
def test_forces_shift_forces_and_energies_match_the_oracle(nb, case):
    pr, cs = _system(**CASES[case])
    nb.setup(cs, pr.params)
    want = nb_oracle.run_port(cs, pr.params, table=None)
    got = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
    _check(got, want, case)
    assert np.max(np.abs(got["fshift"] - want["fshift"])) <= 2e-5 * np.max(np.abs(want["fshift"]))
    assert abs(got["vc"] - want["vc"]) <= 1e-5 * abs(want["vc"])
    assert abs(got["vvdw"] - want["vvdw"]) <= 1e-5 * abs(want["vvdw"])
    assert nb.cluster_pairs == cs.n_cluster_pairs
    # masked and filler atoms receive nothing
    assert not got["f"][cs.perturbed_slots].any() and not got["f"][cs.atom_index < 0].any()


def test_force_only_pass_and_output_semantics(nb):
    pr, cs = _system(**CASES["ewald"])
    nb.setup(cs, pr.params)
    x = cs.xq[:, :3]
    full = nb.compute(x, cs.shiftvec, ALL)
    fo = nb.compute(x, cs.shiftvec, P.DO_FORCE)
    assert fo["vc"] == 0 and fo["vvdw"] == 0 and not fo["fshift"].any()
    # float atomics: the order of the additions differs from launch to launch
    scale = np.max(np.abs(full["f"]))
    assert np.max(np.abs(fo["f"] - full["f"])) <= 2e-6 * scale
    # accumulate (default) and overwrite
    out = dict(f=np.full((cs.natoms, 3), 0.5, np.float32), fshift=np.full((45, 3), 0.25, np.float32), vc=1.0, vvdw=2.0)
    nb.compute(x, cs.shiftvec, ALL, out=out)
    assert np.max(np.abs(out["f"] - 0.5 - full["f"])) <= 4e-6 * scale
    assert abs(out["vc"] - 1.0 - full["vc"]) <= 1e-6 * abs(full["vc"]) and abs(out["vvdw"] - 2.0 - full["vvdw"]) <= 1e-6 * abs(full["vvdw"])
    nb.compute(x, cs.shiftvec, ALL | P.CLEAR_OUTPUTS, out=out)
    assert np.max(np.abs(out["f"] - full["f"])) <= 2e-6 * scale
    assert abs(out["vc"] - full["vc"]) <= 1e-6 * abs(full["vc"])


def test_masking_on_the_device_is_the_reference_s(nb):
    pr, cs = _system(**CASES["ewald"])
    nb.set_params(pr.params)
    nb.set_nbfp(cs.ntype, cs.nbfp)
    nb.set_atoms(cs.type_unmasked, cs.q_unmasked)
    t, q = nb.get_atoms()
    assert np.array_equal(t, cs.type_unmasked) and np.array_equal(q, cs.q_unmasked)
    nb.mask_perturbed(cs.perturbed_slots)
    t, q = nb.get_atoms()
    xq_u = cs.xq.copy()
    xq_u[:, 3] = cs.q_unmasked
    xq_m, t_m = nb_oracle.mask_perturbed(xq_u, cs.type_unmasked, cs.ntype, cs.perturbed_slots)
    assert np.array_equal(t, t_m) and np.array_equal(q, xq_m[:, 3].astype(np.float32))
    # without the mask the perturbed atoms would interact here AND in the FEP kernel
    nb.set_pairlist(cs.sci, cs.cj, cs.excl)
    masked = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
    nb.set_atoms(cs.type_unmasked, cs.q_unmasked)
    nb.set_pairlist(cs.sci, cs.cj, cs.excl)
    unmasked = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
    assert abs(unmasked["vc"] - masked["vc"]) > 1e-3 * abs(masked["vc"])


def test_malformed_input_is_refused(nb):
    from fepb200.lib import FepError

    pr, cs = _system(**CASES["ewald"])
    nb.setup(cs, pr.params)
    bad = cs.cj.copy()
    bad["cj"][5, 2] = cs.natoms  # a j-cluster beyond the atoms
    with pytest.raises(FepError):
        nb.set_pairlist(cs.sci, bad, cs.excl)
    bad = cs.sci.copy()
    bad["shift"][0] = 45
    with pytest.raises(FepError):
        nb.set_pairlist(bad, cs.cj, cs.excl)
    bad = cs.cj.copy()
    bad["excl1"][0] = cs.excl.shape[0]
    with pytest.raises(FepError):
        nb.set_pairlist(cs.sci, bad, cs.excl)
    with pytest.raises(FepError):
        nb.set_atoms(cs.type_unmasked[:-3], cs.q_unmasked[:-3])  # not whole clusters
    nbfp = cs.nbfp.copy()
    nbfp[-1] = 1.0  # the last type must not interact
    with pytest.raises(FepError):
        nb.set_nbfp(cs.ntype, nbfp)
    nb.setup(cs, pr.params)  # still usable
    nb.compute(cs.xq[:, :3], cs.shiftvec, P.DO_FORCE)


def test_empty_list(nb):
    pr, cs = _system(**CASES["ewald"])
    nb.setup(cs, pr.params)
    nb.set_pairlist(cs.sci[:0], cs.cj[:0], cs.excl[:1])
    out = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
    assert not out["f"].any() and out["vc"] == 0 and out["vvdw"] == 0


@pytest.mark.parametrize("case", ["ewald", "rf"])
def test_perturbed_and_non_perturbed_kernels_share_one_device_force_buffer(nb, case):
    """The whole short-range non-bonded force of a system with perturbed atoms, on the device in grid order: the cluster
    kernel (masked atoms) and the perturbed-pair kernels (list in grid indices, original parameters) add into the same
    float3 array; energies of both from device-resident results.  Against the sum of the two CPU checkers."""
    import torch

    from fepb200.lib import FepContext
    from oracle import oracle

    pr, cs = _system(**CASES[case])
    nb.setup(cs, pr.params)
    fl, qA, qB, tA, tB = synth_nb.fep_list_in_slots(pr, cs)
    import copy

    grid = copy.copy(pr)
    grid.nblist, grid.qA, grid.qB, grid.typeA, grid.typeB = fl, qA, qB, tA, tB
    grid.x = np.ascontiguousarray(cs.xq[:, :3])
    fep = FepContext(0)
    try:
        fep.set_problem(grid)
        xq = cs.xq.copy()
        xq[:, 3] = 77.0  # neither kernel may read the charge from here
        d_xq = torch.from_numpy(xq).cuda()
        d_f = torch.zeros((cs.natoms, 3), dtype=torch.float32, device="cuda")
        d_fshift = torch.zeros(135, dtype=torch.float32, device="cuda")
        d_e = torch.zeros(2, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        flags = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL
        nb.launch_device(d_xq.data_ptr(), cs.shiftvec, flags, d_f.data_ptr(), d_fshift.data_ptr(), d_e.data_ptr())
        nb.wait()
        fep.gather_xq_device(d_xq.data_ptr(), cs.shiftvec)
        fep.launch(flags)
        fep.add_forces_device(d_f.data_ptr(), P.ATOMIC_OUTPUTS)
        fep.wait()
        got_fep = fep.download(flags & ~P.DO_FORCE)
        f = d_f.cpu().numpy().astype(np.float64)
    finally:
        fep.close()
    want_nb = nb_oracle.run_port(cs, pr.params, table=None)
    want_fep = oracle.run_port(grid, flags)
    want_f = want_nb["f"] + want_fep["f"]
    rms = np.sqrt(np.mean((f - want_f) ** 2) / np.mean(want_f**2))
    assert rms <= 2e-6
    assert np.abs(want_fep["f"][cs.perturbed_slots]).max() > 1.0  # the perturbed atoms do feel forces, from the FEP side
    e = d_e.cpu().numpy()
    assert abs(e[0] - want_nb["vc"]) <= 1e-5 * abs(want_nb["vc"]) and abs(e[1] - want_nb["vvdw"]) <= 1e-5 * abs(want_nb["vvdw"])
    fsh = d_fshift.cpu().numpy().reshape(45, 3)
    assert np.max(np.abs(fsh - want_nb["fshift"])) <= 2e-5 * np.max(np.abs(want_nb["fshift"]))
    assert np.allclose(got_fep["Vc"], want_fep["Vc"], rtol=1e-4, atol=1e-4 * np.max(np.abs(want_fep["Vc"])))


def test_charges_can_come_from_the_caller_s_xq(nb):
    """FEPB200_NB_Q_FROM_XQ: the fork keeps masked charges in NBAtomDataGpu::xq.w; the kernel then reads the caller's
    array directly (no packing pass)."""
    import torch

    from fepb200.nb import NB_Q_FROM_XQ

    pr, cs = _system(**CASES["ewald"])
    nb.setup(cs, pr.params)
    d_xq = torch.from_numpy(cs.xq).cuda()  # masked charges in .w
    d_f = torch.zeros((cs.natoms, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    n0 = nb.launch_count
    nb.launch_device(d_xq.data_ptr(), cs.shiftvec, P.DO_FORCE | NB_Q_FROM_XQ, d_f.data_ptr())
    nb.wait()
    assert nb.launch_count == n0 + 1
    want = nb_oracle.run_port(cs, pr.params, table=None, energy=False)
    _check(dict(f=d_f.cpu().numpy()), want)
    assert nb.last_kernel_ms() > 0


def test_medium_system_properties(nb):
    """C3-sized (100 k atoms): too slow for the CPU checker in a unit test, so size-independent properties: Newton's third
    law (the forces sum to zero), two launches agree to the rounding of the atomic additions' order, and so do their
    energies."""
    pr = make_system("C3")
    cs = synth_nb.build_cluster_system(pr, rlist=1.1)
    nb.setup(cs, pr.params)
    x = cs.xq[:, :3]
    a = nb.compute(x, cs.shiftvec, ALL)
    b = nb.compute(x, cs.shiftvec, ALL)
    scale = np.max(np.abs(a["f"]))
    assert np.max(np.abs(a["f"].astype(np.float64).sum(axis=0))) <= 1e-4 * scale
    assert np.max(np.abs(a["f"] - b["f"])) <= 4e-6 * scale
    assert abs(a["vc"] - b["vc"]) <= 1e-9 * abs(a["vc"]) + 1e-3
    assert nb.cluster_pairs == cs.n_cluster_pairs > 1_000_000


def test_reference_cuda_kernel_on_the_same_list(nb):
    """The reference's own CUDA cluster-pair kernel (nbnxm_cuda_kernel.cuh compiled in place, oracle/_ref/libnbfork_cuda.so)
    on the same list and the same B200: a second, independent implementation of the reference that must agree with the
    oracle and with us (its analytical Ewald is the reference's own rational fit, ours is another: 5e-6 rel-RMS)."""
    if not nb_oracle.have_fork_cuda():
        pytest.skip("oracle/_ref/libnbfork_cuda.so not built")
    for case in ("ewald", "rf", "ewald_no_overlaps", "rf_no_overlaps"):
        pr, cs = _system(**CASES[case])
        nb.setup(cs, pr.params)
        want = nb_oracle.run_port(cs, pr.params, table=None)
        fork = nb_oracle.run_fork_cuda(cs, pr.params, energy=True, repeats=2)
        ours = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
        for got in (fork, ours):
            rms = np.sqrt(np.mean((np.asarray(got["f"], np.float64) - want["f"]) ** 2) / np.mean(want["f"] ** 2))
            assert rms <= 5e-6, (case, rms)
            assert abs(got["vc"] - want["vc"]) <= 5e-5 * abs(want["vc"]) and abs(got["vvdw"] - want["vvdw"]) <= 5e-5 * abs(want["vvdw"])
        # the CUDA kernel leaves out the shift force of the central cell, which multiplies a zero shift vector in the virial
        # (nbnxm_cuda_kernel.cuh: `if (nb_sci.shift == c_centralShiftIndex) bCalcFshift = false`)
        rows = np.arange(45) != 22
        assert not fork["fshift"][22].any()
        assert np.max(np.abs(fork["fshift"][rows] - want["fshift"][rows])) <= 1e-4 * np.max(np.abs(want["fshift"][rows]))


def test_device_shift_vectors_and_energies_into_float_buffers(nb):
    """What a hook in the fork's GPU route hands over: shift vectors on the device (NBAtomDataGpu::shiftVec), masked charges in
    xq.w, energies added into the fork's float accumulators eLJ / eElec."""
    import torch

    from fepb200.nb import NB_Q_FROM_XQ, NB_SHIFTVEC_ON_DEVICE

    pr, cs = _system(**CASES["ewald"])
    nb.setup(cs, pr.params)
    d_xq = torch.from_numpy(cs.xq).cuda()
    d_sv = torch.from_numpy(np.ascontiguousarray(cs.shiftvec, np.float32)).cuda()
    d_f = torch.zeros((cs.natoms, 3), dtype=torch.float32, device="cuda")
    d_fs = torch.zeros(135, dtype=torch.float32, device="cuda")
    d_e = torch.full((2,), 0.5, dtype=torch.float32, device="cuda")  # {eLJ, eElec}, not cleared by us
    torch.cuda.synchronize()
    flags = ALL | NB_Q_FROM_XQ | NB_SHIFTVEC_ON_DEVICE
    nb.launch_device_raw(d_xq.data_ptr(), d_sv.data_ptr(), flags, d_f.data_ptr(), d_fs.data_ptr(), 0)
    nb.export_energies_device(d_e.data_ptr(), d_e.data_ptr() + 4)
    nb.wait()
    want = nb_oracle.run_port(cs, pr.params, table=None)
    _check(dict(f=d_f.cpu().numpy()), want)
    e = d_e.cpu().numpy().astype(np.float64)
    assert abs(e[0] - 0.5 - want["vvdw"]) <= 1e-5 * abs(want["vvdw"]) and abs(e[1] - 0.5 - want["vc"]) <= 1e-5 * abs(want["vc"])
    fsh = d_fs.cpu().numpy().reshape(45, 3)
    assert np.max(np.abs(fsh - want["fshift"])) <= 2e-5 * np.max(np.abs(want["fshift"]))
    # the same in ONE launch: the kernel adds its energies into the float accumulators itself; three launches in a row use
    # the two work-queue heads in turn
    for _ in range(3):
        d_f.zero_()
        d_fs.zero_()
        d_e.fill_(0.5)
        torch.cuda.synchronize()
        n0 = nb.launch_count
        nb.launch_device_float_energies(d_xq.data_ptr(), d_sv.data_ptr(), flags, d_f.data_ptr(), d_fs.data_ptr(), d_e.data_ptr(),
                                        d_e.data_ptr() + 4)
        nb.wait()
        assert nb.launch_count == n0 + 1
        _check(dict(f=d_f.cpu().numpy()), want)
        e = d_e.cpu().numpy().astype(np.float64)
        assert abs(e[0] - 0.5 - want["vvdw"]) <= 2e-5 * abs(want["vvdw"]) and abs(e[1] - 0.5 - want["vc"]) <= 2e-5 * abs(want["vc"])


def test_list_read_from_the_caller_s_device_copy_with_pruned_masks(nb):
    """fepb200_nb_use_device_list: the kernel reads the caller's device copy of the list (the fork's gpu_plist) -- work items from
    the host list, i-cluster masks from the device copy at kernel time.  The fork's pruning kernels clear mask bits there, per
    HALF of a cluster pair (each of its two warps prunes its own half): the kernel must take the union of the two halves."""
    import copy

    import torch

    from fepb200.nb import NB_Q_FROM_XQ

    pr, cs = _system(**CASES["ewald"])
    nb.setup(cs, pr.params)
    rng = np.random.default_rng(3)
    cj = cs.cj.copy()
    n = cj.shape[0]
    pick = rng.random(n)
    drop = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    a = pick < 0.3  # one half pruned away, the other kept: nothing may change for these
    cj["imask0"][a] &= drop[a]
    b = (pick >= 0.3) & (pick < 0.6)  # both halves pruned alike: these cluster pairs are gone
    cj["imask0"][b] &= drop[b]
    cj["imask1"][b] &= drop[b]
    pruned = copy.copy(cs)
    pruned.cj = cs.cj.copy()
    pruned.cj["imask0"] = cj["imask0"] | cj["imask1"]
    pruned.cj["imask1"] = pruned.cj["imask0"]
    want = nb_oracle.run_port(pruned, pr.params, table=None, energy=False)
    full = nb_oracle.run_port(cs, pr.params, table=None, energy=False)
    assert np.max(np.abs(want["f"] - full["f"])) > 1e-3 * np.max(np.abs(full["f"]))  # the pruning does remove interactions
    d_sci = torch.from_numpy(cs.sci.view(np.int32).reshape(-1, 4).copy()).cuda()
    d_cj = torch.from_numpy(cj.view(np.int32).reshape(-1, 8).copy()).cuda()
    d_excl = torch.from_numpy(cs.excl.view(np.int32).reshape(-1, 32).copy()).cuda()
    d_xq = torch.from_numpy(cs.xq).cuda()
    d_f = torch.zeros((cs.natoms, 3), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    nb.use_device_list(d_sci.data_ptr(), d_cj.data_ptr(), d_excl.data_ptr())
    nb.launch_device(d_xq.data_ptr(), cs.shiftvec, P.DO_FORCE | NB_Q_FROM_XQ, d_f.data_ptr())
    nb.wait()
    # a random half of the atom pairs of a system with close contacts: the forces no longer cancel and the steep close
    # pairs (r of 0.05 nm from fp32 coordinates of 4 nm) set the deviation: 4.5e-6 measured, the budget of the path is 1e-5
    _check(dict(f=d_f.cpu().numpy()), want, "pruned device list", core_tol=1e-5)
    # back to the library's own copy of the host list
    nb.use_device_list(0, 0, 0)
    d_f.zero_()
    torch.cuda.synchronize()
    nb.launch_device(d_xq.data_ptr(), cs.shiftvec, P.DO_FORCE | NB_Q_FROM_XQ, d_f.data_ptr())
    nb.wait()
    _check(dict(f=d_f.cpu().numpy()), full, "own list")


@pytest.mark.parametrize("modifier", ["forceswitch", "potswitch"])
def test_lennard_jones_switch_flavours_of_the_reference_s_cuda_kernels(nb, modifier):
    """vdw_modifier = force switch / potential switch: flavours of the reference's CUDA kernels (nbnxm_cuda_kernel_utils.cuh:104-211)
    that nbnxn_kernel_gpu_ref does not have.  The checker applies them the way those kernels do (oracle/nb_oracle.c,
    cuda_modifiers); it is pinned HERE, on the GPU, against the reference's own kernels nbnxn_kernel_ElecEw_VdwLJ{Fsw,Psw}_{F,VF}_cuda
    compiled in place, and then compared with ours."""
    import copy

    # without overlapping waters: with them one pair's 1e10 kJ/mol would hide what the switch does to the sums
    pr, cs = _system(**CASES["ewald_no_overlaps"])
    base = P.make_params(coulombtype="pme", vdw_modifier=modifier, rvdw_switch=0.8)
    params = copy.copy(pr.params)
    for k in ("vdw_modifier", "rvdw_switch", "dispersion_shift_cpot", "repulsion_shift_cpot"):
        setattr(params, k, getattr(base, k))
    params = params.rounded()
    plain = nb_oracle.run_port(cs, params, table=None)  # what nbnxn_kernel_gpu_ref would give: the switch ignored
    want = nb_oracle.run_port(cs, params, table=None, cuda_modifiers=True)
    mag = np.linalg.norm(plain["f"], axis=1)
    assert np.max(np.abs(want["f"] - plain["f"])) > 1e-4 * np.median(mag[mag > 0])  # the switch is felt
    assert abs(want["vvdw"] - plain["vvdw"]) > 1e-2 * abs(plain["vvdw"])
    if nb_oracle.have_fork_cuda():
        fork = nb_oracle.run_fork_cuda(cs, params, energy=True, repeats=1)
        rms = np.sqrt(np.mean((fork["f"] - want["f"]) ** 2) / np.mean(want["f"] ** 2))
        assert rms <= 5e-6, rms
        assert abs(fork["vvdw"] - want["vvdw"]) <= 5e-5 * abs(want["vvdw"]) and abs(fork["vc"] - want["vc"]) <= 5e-5 * abs(want["vc"])
    nb.setup(cs, params)
    got = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
    _check(got, want, modifier)
    assert abs(got["vvdw"] - want["vvdw"]) <= 1e-5 * abs(want["vvdw"]) and abs(got["vc"] - want["vc"]) <= 1e-5 * abs(want["vc"])
    fo = nb.compute(cs.xq[:, :3], cs.shiftvec, P.DO_FORCE)
    assert np.max(np.abs(fo["f"] - got["f"])) <= 4e-6 * np.max(np.abs(got["f"]))
    nb.setup(cs, pr.params)  # back to the plain flavour for the tests that follow
