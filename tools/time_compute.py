"""Wall time of fepb200_compute() (host buffers in, host buffers out), with the library's own phase
laps when FEPB200_TIMING is set.  Other knobs are read by the library from the environment."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gromacs-fep-gpu_b200', 'python'), ROOT]
import numpy as np
from fepb200 import params as P
from fepb200.lib import FepContext
from fepb200.synth import make_system
ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
prob = make_system(name)
with FepContext(0) as ctx:
    ctx.set_problem(prob)
    out = ctx.new_outputs()
    x = np.ascontiguousarray(prob.x); sv = np.ascontiguousarray(prob.shiftvec)
    ref = ctx.new_outputs()
    ctx.compute(x, sv, ALL | P.CLEAR_OUTPUTS, ref)
    for _ in range(10):
        ctx.compute(x, sv, ALL | P.CLEAR_OUTPUTS, out)
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.compute(x, sv, ALL | P.CLEAR_OUTPUTS, out)
    dt = (time.perf_counter() - t0) / n
    same = all(np.array_equal(np.asarray(out[k]), np.asarray(ref[k])) for k in ("f", "fshift", "Vc", "Vv", "dvdl", "foreign_energy"))
    env = " ".join(f"{k[8:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FEPB200_") and k != "FEPB200_TIMING")
    print(f"{name} [{env}] compute() {dt*1e6:.1f} us  (repeatable: {same})")
