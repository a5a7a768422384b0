"""Pins oracle/nb_oracle.c -- the CPU checker of the non-perturbed cluster-pair kernel -- before any GPU
result is compared with it:

  1. against oracle/_ref/libnbref_dp.so = the reference's own kernels_reference/kernel_gpu_ref.cpp compiled in place,
     double precision, same list, same Ewald force table: agreement at rounding level;
  2. list structures: byte sizes and constants of the reference's nbnxn_sci_t / nbnxn_cj_packed_t / nbnxn_excl_t against
     the numpy dtypes of the input generator and include/fepb200_nb.h;
  3. list generator + kernel semantics against an all-pairs (minimum-image) sum with the same exclusions;
  4. the analytical Ewald force (what the GPU kernel evaluates) against the tabulated one the reference interpolates;
  5. masked perturbed atoms (atomdata.cpp:930-964): they contribute nothing, their exclusion bits do not matter.
No GPU involved.
"""
import numpy as np
import pytest

from fepb200 import synth_nb
from fepb200.synth import make_system, scaled_spec
from oracle import nb_oracle

needs_ref = pytest.mark.skipif(not nb_oracle.have_ref("dp"), reason="oracle/_ref/libnbref_dp.so not built")


def _system(name, box, n_blobs, seed=5, rlist=1.1, **kw):
    split = kw.pop("max_cj_groups_per_sci", 0)
    pr = make_system(scaled_spec(name, box, n_blobs, **kw), seed=seed)
    return pr, synth_nb.build_cluster_system(pr, rlist=rlist, max_cj_groups_per_sci=split)


CASES = {
    "ewald": dict(name="C2", box=4.2, n_blobs=1),
    "rf": dict(name="C4", box=4.2, n_blobs=1, n_energy_groups=1),
    "ewald_split_entries": dict(name="C1", box=3.6, n_blobs=1, max_cj_groups_per_sci=3),
}


def _rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-30)


@needs_ref
def test_struct_layout_is_the_reference_s():
    sizes = nb_oracle.ref_struct_sizes("dp")
    assert sizes == [synth_nb.SCI_DTYPE.itemsize, synth_nb.CJ_DTYPE.itemsize, synth_nb.EXCL_DTYPE.itemsize,
                     synth_nb.CL, synth_nb.NCL_SC, synth_nb.JGROUP, 2]
    assert nb_oracle.ref_struct_sizes("sp")[:3] == sizes[:3]


@needs_ref
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("energy", [True, False])
def test_port_reproduces_the_reference_kernel(case, energy):
    pr, cs = _system(**CASES[case])
    table = (2000.0, 4096)
    ref = nb_oracle.run_ref(cs, pr.params, energy=energy, table=table, precision="dp")
    port = nb_oracle.run_port(cs, pr.params, energy=energy, table=table, min_rsq=nb_oracle.MIN_RSQ_DOUBLE)
    assert np.max(np.abs(ref["f"])) > 10.0
    assert _rel(port["f"], ref["f"]) < 1e-11
    assert _rel(port["fshift"], ref["fshift"]) < 1e-10
    if energy:
        assert abs(port["vc"] - ref["vc"]) < 1e-10 * abs(ref["vc"])
        assert abs(port["vvdw"] - ref["vvdw"]) < 1e-10 * abs(ref["vvdw"])
    else:
        assert ref["vc"] == 0 and ref["vvdw"] == 0 and port["vc"] == 0 and port["vvdw"] == 0


@needs_ref
def test_float_build_of_the_reference_sets_the_error_budget():
    """The mixed-precision build of the reference kernel against its double build: what fp32 pair maths costs."""
    pr, cs = _system(**CASES["ewald"])
    dp = nb_oracle.run_ref(cs, pr.params, precision="dp")
    sp = nb_oracle.run_ref(cs, pr.params, precision="sp")
    rms = np.sqrt(np.mean((sp["f"] - dp["f"]) ** 2) / np.mean(dp["f"] ** 2))
    assert rms < 2e-6
    assert abs(sp["vc"] - dp["vc"]) < 2e-5 * abs(dp["vc"])


@pytest.mark.parametrize("case", ["ewald", "rf"])
def test_list_and_kernel_against_all_pairs(case):
    pr, cs = _system(**CASES[case])
    box = float(cs.shiftvec[23, 0])  # the float32 box edge the shift vectors carry
    # the exclusions the generator applied, as slot pairs
    n_sol = pr.perturbed.size
    w0 = np.arange(n_sol, pr.natoms, 3)
    fl = pr.nblist
    ex = np.concatenate([np.stack([w0, w0 + 1], 1), np.stack([w0, w0 + 2], 1), np.stack([w0 + 1, w0 + 2], 1),
                         np.stack([np.repeat(fl.iinr, np.diff(fl.jindex)), fl.jjnr], 1)])
    ex = ex[ex[:, 0] != ex[:, 1]]
    exs = cs.slot_of_atom[ex]
    want_f, want_vc, want_vv = synth_nb.brute_force(cs, pr.params, exs, box, pr.params.elec_ewald)
    got = nb_oracle.run_port(cs, pr.params, table=None)
    assert _rel(got["f"], want_f) < 1e-9
    assert abs(got["vc"] - want_vc) < 1e-9 * abs(want_vc)
    assert abs(got["vvdw"] - want_vv) < 1e-9 * abs(want_vv)
    # shift forces sum to the total force on the i side: with every image present they cancel the net force
    assert np.max(np.abs(got["f"].sum(axis=0))) < 1e-6 * np.max(np.abs(got["f"]))


def test_analytical_ewald_force_against_the_table():
    pr, cs = _system(**CASES["ewald"])
    tab = nb_oracle.run_port(cs, pr.params, table=(2000.0, 4096))
    ana = nb_oracle.run_port(cs, pr.params, table=None)
    rms = np.sqrt(np.mean((tab["f"] - ana["f"]) ** 2) / np.mean(ana["f"] ** 2))
    assert 0 < rms < 1e-6  # linear interpolation at 2000 points/nm
    assert tab["vc"] == ana["vc"]  # the energy never uses the table (kernel_gpu_ref.cpp:250-256)


def test_masked_atoms_do_not_interact():
    pr, cs = _system(**CASES["ewald"])
    # masking on the unmasked arrays gives the arrays the generator made
    xq_u = cs.xq.copy()
    xq_u[:, 3] = cs.q_unmasked
    xq_m, type_m = nb_oracle.mask_perturbed(xq_u, cs.type_unmasked, cs.ntype, cs.perturbed_slots)
    assert np.array_equal(xq_m.astype(np.float32), cs.xq) and np.array_equal(type_m, cs.type)
    assert np.any(cs.q_unmasked[cs.perturbed_slots] != 0)
    base = nb_oracle.run_port(cs, pr.params)
    assert np.all(base["f"][cs.perturbed_slots] == 0)
    assert np.all(base["f"][cs.atom_index < 0] == 0)
    # the interaction bits of pairs with a masked atom are irrelevant
    # (dropping ALL exclusions would change the water molecules: rebuild the list without the FEP pairs' bits instead)
    pr3 = make_system(scaled_spec("C2", 4.2, 1), seed=5)
    pr3.nblist = pr3.nblist.select_pairs(np.zeros(pr3.nblist.nrj, bool))  # no FEP pairs -> no bits cleared for them
    cs3 = synth_nb.build_cluster_system(pr3, rlist=1.1)
    alt = nb_oracle.run_port(cs3, pr.params)
    assert np.array_equal(cs3.atom_index, cs.atom_index)
    assert _rel(alt["f"], base["f"]) < 1e-12 and abs(alt["vc"] - base["vc"]) < 1e-12 * abs(base["vc"])
