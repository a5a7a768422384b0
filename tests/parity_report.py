"""Prints the measured deviation of the GPU path from the double-precision oracle for the five
BASELINE.json configurations at full size (run on the GPU box; output committed under profiles/).  Test infrastructure: it lives in tests/ because it calls the oracle.."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT]
import numpy as np  # noqa: E402

from fepb200 import params as P  # noqa: E402
from fepb200.lib import FepContext  # noqa: E402
from fepb200.synth import SPECS, make_system  # noqa: E402
from oracle import oracle  # noqa: E402

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.maximum(np.abs(b), 1e-2 * np.max(np.abs(b)))
    return float(np.max(np.abs(a - b) / np.where(scale > 0, scale, 1.0)))


def force_rms(f, ref, keep=None):
    f, ref = np.asarray(f, float), np.asarray(ref, float)
    if keep is not None:
        f, ref = f[keep], ref[keep]
    return float(np.sqrt(np.mean((f - ref) ** 2) / np.mean(ref**2)))


def ordinary_atoms(ref):
    """Atoms that take part and whose force is not astronomically larger than the others' (|f| <= 100 x the 99th percentile):
    the adversarial placements of the synthetic systems can put a water next to a hard-cored perturbed atom (C2, C5), and
    one such pair (1e8 kJ/mol/nm) then is the whole of an RMS over all atoms."""
    mag = np.linalg.norm(np.asarray(ref, float), axis=1)
    on = mag > 0
    return on & (mag <= 100.0 * np.percentile(mag[on], 99)), int(on.sum()), float(mag.max()), float(np.percentile(mag[on], 99))


print(f"{'cfg':4s} {'atoms':>8s} {'pairs':>8s} {'L':>3s} {'force rel-RMS':>14s} {'Vc':>9s} {'Vv':>9s} {'dvdl':>9s} "
      f"{'foreign E':>10s} {'foreign dvdl':>12s}   oracle   (budget: 1e-5 forces, 1e-4 the rest)")
with FepContext(0) as ctx:
    for name in ("C1", "C2", "C3", "C4", "C5"):
        prob = make_system(SPECS[name])
        ctx.set_problem(prob)
        out = ctx.compute(prob.x, prob.shiftvec, ALL)
        ref = oracle.run_best(prob, ALL, nthreads=min(16, os.cpu_count() or 1))
        rms = np.sqrt(np.mean((out["f"] - ref["f"]) ** 2) / np.mean(ref["f"] ** 2))
        keep, n_on, f_max, f_p99 = ordinary_atoms(ref["f"])
        trimmed = (f"   [forces of the {int(keep.sum())} of {n_on} atoms with |f| <= 100 x p99 (p99 {f_p99:.3g}, max {f_max:.3g}): "
                   f"rel-RMS {force_rms(out['f'], ref['f'], keep):.2e}")
        sp = ""
        if oracle.have_ref("sp"):
            r32 = oracle.run_ref(prob, ALL, precision="sp", nthreads=min(16, os.cpu_count() or 1))
            rms32 = np.sqrt(np.mean((r32["f"] - ref["f"]) ** 2) / np.mean(ref["f"] ** 2))
            trimmed += f", the reference's fp32 build {force_rms(r32['f'], ref['f'], keep):.2e}"
            sp = f"   [reference fp32 build vs fp64: force {rms32:.1e}, foreign E {rel(r32['foreign_energy'], ref['foreign_energy']):.1e}, dvdl {rel(r32['dvdl'], ref['dvdl']):.1e}]"
        print(f"{name:4s} {prob.natoms:8d} {prob.nblist.nrj:8d} {prob.n_foreign:3d} {rms:14.2e} {rel(out['Vc'], ref['Vc']):9.1e} "
              f"{rel(out['Vv'], ref['Vv']):9.1e} {rel(out['dvdl'], ref['dvdl']):9.1e} "
              f"{rel(out['foreign_energy'], ref['foreign_energy']):10.1e} {rel(out['foreign_dvdl'], ref['foreign_dvdl']):12.1e}   "
              f"{ref['variant']}{sp}{trimmed}]", flush=True)
