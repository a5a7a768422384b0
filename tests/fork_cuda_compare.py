"""The reference fork's own CUDA FEP kernels (oracle/_ref/libfepfork_cuda.so, compiled in place for
sm_100a) beside ours on the same B200 and the same problems.  Test / measurement infrastructure:
it lives in tests/ because it loads libraries from oracle/.

  python tests/fork_cuda_compare.py small            # two small problems, with the fp64 oracle (seconds)
  python tests/fork_cuda_compare.py C5 [C2 C4g1 ...] # full-size BASELINE configurations, timing + cross-check

Prints one JSON line per problem: device times of the fork's kernels (current-lambda kernel, foreign
kernel; CUDA events, best of 20, warm caches) and of ours (pass / foreign / epilogue kernels from the
library's own events, and the whole step), and the deviations fork-vs-ours (and both vs the oracle on
the small problems).  The fork has no energy groups, no Gapsys, clamps r^2 at 3.8e-7 instead of 1e-12
and does not test the soft-core radius against the cut-off, so problems are generated
without adversarial placements and with one energy group; deviations are reported, the caller judges.
"""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT]
import numpy as np  # noqa: E402

from fepb200 import params as P  # noqa: E402
from fepb200.lib import FepContext  # noqa: E402
from fepb200.synth import SPECS, make_system, scaled_spec  # noqa: E402
from oracle import oracle  # noqa: E402

ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA


def rel_rms(a, b):
    b = np.asarray(b, float)
    return float(np.sqrt(np.mean((np.asarray(a, float) - b) ** 2) / max(np.mean(b**2), 1e-300)))


def rel_max(a, b):
    b = np.asarray(b, float)
    return float(np.max(np.abs(np.asarray(a, float) - b)) / max(np.max(np.abs(b)), 1e-300))


def specs(names):
    for name in names:
        if name == "small":
            yield "small-C2", scaled_spec("C2", 3.6, 1, 30, n_foreign=6, n_adversarial=0), True
            yield "small-C4g1", scaled_spec("C4", 4.2, 2, 25, n_foreign=8, n_energy_groups=1, n_adversarial=0), True
        elif name == "C4g1":  # C4 with one energy group (the fork has none)
            yield name, dataclasses.replace(SPECS["C4"], n_energy_groups=1, n_adversarial=0), False
        else:
            yield name, dataclasses.replace(SPECS[name], n_adversarial=0), False


def main():
    names = sys.argv[1:] or ["small"]
    if not oracle.have_fork_cuda():
        print(json.dumps(dict(unavailable="oracle/_ref/libfepfork_cuda.so not built (make -C oracle fork_cuda)")))
        return
    with FepContext(0) as ctx:
        for label, spec, with_oracle in specs(names):
            prob = make_system(spec)
            why = oracle.fork_cuda_unsupported(prob)
            if why:
                print(json.dumps(dict(problem=label, unavailable="fork GPU kernels do not cover: " + why)), flush=True)
                continue
            ctx.set_problem(prob)
            ours = ctx.compute(prob.x, prob.shiftvec, ALL)
            # our device times: resident inputs, per-kernel events of the library + whole step
            ctx.upload_x(prob.x, prob.shiftvec)
            step_ms, kms = [], []
            for profiling in (False, True):
                ctx.set_profiling(profiling)
                for _ in range(20):
                    ctx.launch(ALL)
                    ctx.wait()
                    (kms if profiling else step_ms).append(ctx.kernel_ms() if profiling else ctx.last_launch_ms())
            ctx.set_profiling(False)
            try:
                fork = oracle.run_fork_cuda(prob, ALL, repeats=20)
            except Exception as exc:  # noqa: BLE001 -- foreign kernels in this process: a device fault of theirs is sticky, stop here
                print(json.dumps(dict(problem=label, fork_error=f"{type(exc).__name__}: {exc}"[:800])), flush=True)
                return
            line = dict(
                problem=label, natoms=int(prob.natoms), pairs=int(prob.nblist.nrj), entries=int(prob.nblist.nri),
                n_foreign=int(prob.n_foreign),
                fork_us=dict(current_lambda_kernel=1e6 * fork["seconds"][0], foreign_kernel=1e6 * fork["seconds"][1],
                             both=1e6 * sum(fork["seconds"])),
                ours_us=dict(step=1e3 * min(step_ms), pass_kernel=1e3 * min(k[0] for k in kms),
                             foreign_kernel=1e3 * min(k[1] for k in kms), epilogue_kernel=1e3 * min(k[2] for k in kms)),
                fork_vs_ours=dict(force_rel_rms=rel_rms(fork["f"], ours["f"]), Vc=rel_max(fork["Vc"], ours["Vc"]),
                                  Vv=rel_max(fork["Vv"], ours["Vv"]), dvdl=rel_max(fork["dvdl"], ours["dvdl"]),
                                  foreign_energy=rel_max(fork["foreign_energy"], ours["foreign_energy"])),
            )
            line["speedup_kernels"] = line["fork_us"]["both"] / max(line["ours_us"]["step"], 1e-9)
            if with_oracle:
                ref = oracle.run_best(prob, ALL)
                for who, res in (("fork_vs_oracle", fork), ("ours_vs_oracle", ours)):
                    line[who] = dict(force_rel_rms=rel_rms(res["f"], ref["f"]), Vc=rel_max(res["Vc"], ref["Vc"]),
                                     Vv=rel_max(res["Vv"], ref["Vv"]), dvdl=rel_max(res["dvdl"], ref["dvdl"]),
                                     foreign_energy=rel_max(res["foreign_energy"], ref["foreign_energy"]))
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
