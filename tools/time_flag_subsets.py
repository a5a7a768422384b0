"""Device time of the step's kernels for subsets of the output flags (which part of the epilogue is
on the critical path?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "gromacs-fep-gpu_b200", "python"), ROOT]
import numpy as np, torch
from fepb200 import params as P
from fepb200.lib import FepContext
from fepb200.synth import make_system

for name in sys.argv[1:] or ("C5",):
    prob = make_system(name)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    with FepContext(0) as ctx:
        ctx.set_stream(stream.cuda_stream)
        ctx.set_problem(prob)
        ctx.upload_x(np.ascontiguousarray(prob.x), prob.shiftvec)
        ctx.set_profiling(True)
        for label, fl in (("F", P.DO_FORCE), ("F+S", P.DO_FORCE | P.DO_SHIFTFORCE), ("F+V", P.DO_FORCE | P.DO_POTENTIAL),
                          ("F+S+V", P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL),
                          ("F+S+V+L", P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA),
                          ("V+L", P.DO_POTENTIAL | P.DO_FOREIGNLAMBDA)):
            k, tot = [], []
            for i in range(30):
                flush.zero_()
                ctx.launch(fl)
                torch.cuda.synchronize()
                if i >= 5:
                    k.append(ctx.kernel_ms()); tot.append(ctx.last_launch_ms())
            k = np.array(k) * 1e3
            print(f"{name} {label:8s} step {np.mean(tot)*1e3:6.1f} us | pass {k[:,0].mean():5.1f} foreign {k[:,1].mean():5.1f} epilogue {k[:,2].mean():5.1f}")
