"""Wall-clock breakdown of one fepb200_compute() call into its phases (host gather + H2D, kernels,
D2H + host scatter) on the bench configurations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gromacs-fep-gpu_b200', 'python'), ROOT]
import numpy as np
from fepb200 import params as P
from fepb200.lib import FepContext
from fepb200.synth import make_system
ALL = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA
for name in sys.argv[1:] or ("C2", "C5"):
    prob = make_system(name)
    with FepContext(0) as ctx:
        ctx.set_problem(prob)
        out = ctx.new_outputs()
        x = np.ascontiguousarray(prob.x); sv = np.ascontiguousarray(prob.shiftvec)
        for _ in range(5):
            ctx.compute(x, sv, ALL | P.CLEAR_OUTPUTS, out)
        n = 200
        t = [0.0] * 4
        for _ in range(n):
            t0 = time.perf_counter(); ctx.upload_x(x, sv); ctx.wait()
            t1 = time.perf_counter(); ctx.launch(ALL); ctx.wait()
            t2 = time.perf_counter(); ctx.download(ALL | P.CLEAR_OUTPUTS, out)
            t3 = time.perf_counter()
            t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2
        t0 = time.perf_counter()
        for _ in range(n):
            ctx.compute(x, sv, ALL | P.CLEAR_OUTPUTS, out)
        t[3] = time.perf_counter() - t0
        lay = ctx.layout()
        print(f"{name}: touched {lay.ntouched}  upload(gather+H2D+sync) {t[0]/n*1e6:.1f} us  launch+sync {t[1]/n*1e6:.1f} us  "
              f"download(D2H+scatter) {t[2]/n*1e6:.1f} us  | compute() {t[3]/n*1e6:.1f} us")
