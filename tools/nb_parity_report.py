"""Deviation of the cluster-pair kernel (csrc/nb/fep_nb.cu) from its fp64 checker (oracle/nb_oracle.c), case by case, on ALL
atoms and on the atoms with ordinary forces.  The synthetic systems carry adversarial placements (fepb200.synth: a water molecule
moved next to a soft-cored atom), which can put two waters on top of each other: one such pair (1e12 kJ/mol/nm) dominates any RMS
and any energy sum, so the report also gives the numbers without them (n_adversarial = 0) and restricted to atoms whose force is
below 50 x the median.  Needs a GPU.   python tools/nb_parity_report.py"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT]
import numpy as np

from fepb200 import params as P
from fepb200 import synth_nb
from fepb200.nb import NbContext
from fepb200.synth import make_system, scaled_spec
from oracle import nb_oracle

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL
CASES = {
    "ewald": dict(name="C2", box=4.2, n_blobs=1),
    "rf": dict(name="C4", box=4.2, n_blobs=1, n_energy_groups=1),
    "ewald_split_entries": dict(name="C1", box=3.6, n_blobs=1, split=3),
    "ewald_larger": dict(name="C3", box=6.0, n_blobs=2),
}


def system(name, box, n_blobs, seed=5, split=0, **kw):
    pr = make_system(scaled_spec(name, box, n_blobs, **kw), seed=seed)
    return pr, synth_nb.build_cluster_system(pr, rlist=1.1, max_cj_groups_per_sci=split)


def dev(got, want):
    f, fw = np.asarray(got["f"], np.float64), want["f"]
    mag = np.linalg.norm(fw, axis=1)
    typical = np.median(mag[mag > 0])
    core = mag <= 50 * typical
    rms = np.sqrt(np.mean((f - fw) ** 2) / np.mean(fw**2))
    rms_core = np.sqrt(np.mean((f[core] - fw[core]) ** 2) / np.mean(fw[core] ** 2))
    out = dict(max_f=float(mag.max()), typical_f=float(typical), outliers=int((~core).sum()), rms=float(rms), rms_core=float(rms_core),
               max_dev_core_over_typical=float(np.max(np.abs(f[core] - fw[core])) / typical))
    for k in ("vc", "vvdw"):
        if k in got and k in want:
            out[k + "_rel"] = float(abs(got[k] - want[k]) / abs(want[k]))
            out[k] = float(want[k])
    if "fshift" in got and "fshift" in want:
        out["fshift_rel_max"] = float(np.max(np.abs(got["fshift"] - want["fshift"])) / np.max(np.abs(want["fshift"])))
    return out


def main():
    nb = NbContext(0)
    for adv in (None, 0):
        for case, kw in CASES.items():
            kw = dict(kw)
            if adv is not None:
                kw["n_adversarial"] = adv
            pr, cs = system(**kw)
            nb.setup(cs, pr.params)
            want = nb_oracle.run_port(cs, pr.params, table=None)
            got = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
            print(f"{case} n_adversarial={'spec' if adv is None else adv}: ours vs oracle", dev(got, want), flush=True)
            if nb_oracle.have_fork_cuda() and case == "ewald":
                fork = nb_oracle.run_fork_cuda(cs, pr.params, energy=True, repeats=1)
                print(f"{case} n_adversarial={'spec' if adv is None else adv}: reference CUDA kernel vs oracle", dev(fork, want), flush=True)
    for modifier in ("forceswitch", "potswitch"):
        pr, cs = system(**CASES["ewald"], n_adversarial=0)
        base = P.make_params(coulombtype="pme", vdw_modifier=modifier, rvdw_switch=0.8)
        params = copy.copy(pr.params)
        for k in ("vdw_modifier", "rvdw_switch", "dispersion_shift_cpot", "repulsion_shift_cpot"):
            setattr(params, k, getattr(base, k))
        params = params.rounded()
        plain = nb_oracle.run_port(cs, params, table=None)
        want = nb_oracle.run_port(cs, params, table=None, cuda_modifiers=True)
        print(f"{modifier}: the switch changes", dev(dict(f=want["f"], vc=want["vc"], vvdw=want["vvdw"]), plain), flush=True)
        if nb_oracle.have_fork_cuda():
            fork = nb_oracle.run_fork_cuda(cs, params, energy=True, repeats=1)
            print(f"{modifier}: reference CUDA kernel vs oracle", dev(fork, want), flush=True)
        nb.setup(cs, params)
        got = nb.compute(cs.xq[:, :3], cs.shiftvec, ALL)
        print(f"{modifier}: ours vs oracle", dev(got, want), flush=True)
    nb.close()


if __name__ == "__main__":
    main()
