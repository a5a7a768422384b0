"""A few device-resident steps of one configuration, for ncu and for quick A/B timing: coordinates are uploaded
once, then `steps` launches (pass + foreign passes + epilogue), optionally with the L2 flushed in between.
  python tools/prof_step.py C5 [steps] [flags: all|force|nofor] [flush] [nf=<foreign lambda points>]
Prints per-step device time (CUDA events of the library around the whole launch) and per-kernel times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gromacs-fep-gpu_b200', 'python'), ROOT]
import numpy as np
import torch
from fepb200 import params as P
from fepb200.lib import FepContext
from fepb200.synth import make_system

name = sys.argv[1] if len(sys.argv) > 1 else "C5"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
what = sys.argv[3] if len(sys.argv) > 3 else "all"
flush = "flush" in sys.argv[4:]
nf = [int(a[3:]) for a in sys.argv[4:] if a.startswith("nf=")]
flags = {"all": P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA,
         "force": P.DO_FORCE,
         "nofor": P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL}[what]
if nf:
    import dataclasses

    from fepb200.synth import SPECS

    prob = make_system(dataclasses.replace(SPECS[name], n_foreign=nf[0]))
else:
    prob = make_system(name)
scratch = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if flush else None
with FepContext(0) as ctx:
    ctx.set_problem(prob)
    ctx.upload_x(prob.x, prob.shiftvec)
    for _ in range(3):
        ctx.launch(flags)
    ctx.wait()
    ms = []
    for _ in range(steps):
        if flush:
            scratch.fill_(1)
            torch.cuda.synchronize()
        ctx.launch(flags)
        ctx.wait()
        ms.append(ctx.last_launch_ms())
    ctx.set_profiling(True)
    kms = []
    for _ in range(steps):
        if flush:
            scratch.fill_(1)
            torch.cuda.synchronize()
        ctx.launch(flags)
        ctx.wait()
        kms.append(ctx.kernel_ms())
    ctx.set_profiling(False)
    k = np.mean(np.array(kms), axis=0) * 1e3
    env = " ".join(f"{k_[8:]}={v}" for k_, v in sorted(os.environ.items()) if k_.startswith("FEPB200_"))
    print(f"{name} {what}{' flush' if flush else ''}{' nf=%d' % nf[0] if nf else ''} [{env}] step {np.median(ms)*1e3:.1f} us (min {np.min(ms)*1e3:.1f}); "
          f"alone: pass {k[0]:.1f} foreign {k[1]:.1f} epilogue {k[2]:.1f} us; pairs {prob.nblist.nrj}")
