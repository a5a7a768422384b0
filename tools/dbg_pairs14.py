import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gromacs-fep-gpu_b200', 'python'), ROOT, os.path.join(ROOT, 'tests')]
import numpy as np
from fepb200 import params as P
from fepb200.pairs14 import Pairs14Context
from pairs14_cases import random_pairs14
from oracle import oracle
FLAGS = P.DO_FORCE | P.DO_SHIFTFORCE | P.DO_POTENTIAL
prob = random_pairs14(8, "beutler")
with Pairs14Context(0) as ctx:
    ctx.set_problem(prob)
    one = ctx.compute(prob, FLAGS)
    two = ctx.compute(prob, FLAGS, out={k: v.copy() for k, v in one.items()})
ref = oracle.run_pairs14(prob)
print("Vv", one["Vv"], ref["Vv"]); print("Vc", one["Vc"], ref["Vc"]); print("dvdl", one["dvdl"], ref["dvdl"])
print("f rms", np.sqrt(np.mean((one["f"]-ref["f"])**2)/np.mean(ref["f"]**2)), "max|f|", np.abs(ref["f"]).max())
d = np.abs(two["f"] - 2*one["f"]); print("accumulate max diff", d.max(), "at", np.unravel_index(d.argmax(), d.shape), two["f"].ravel()[d.argmax()], one["f"].ravel()[d.argmax()])
# per pair distances
ia = prob.iatoms
dx = prob.x[ia[:,1]] - prob.x[ia[:,2]]
dx -= np.rint(dx / prob.box_diag) * prob.box_diag
print("min r", np.sort(np.linalg.norm(dx, axis=1))[:6])
