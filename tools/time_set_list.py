"""Times the search-step entry points (fepb200_set_atoms / fepb200_set_list) on the bench configurations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gromacs-fep-gpu_b200', 'python'), ROOT]
import numpy as np
from fepb200 import params as P
from fepb200.lib import FepContext
from fepb200.synth import make_system
for name in ("C2","C5"):
    prob=make_system(name)
    with FepContext(0) as ctx:
        ctx.set_params(prob.params); ctx.set_nbfp(prob.ntype, prob.nbfp, prob.nbfp_grid)
        t=time.perf_counter(); ctx.set_atoms(prob.qA,prob.qB,prob.typeA,prob.typeB); ta=time.perf_counter()-t
        ctx.set_lambdas(prob.lambda_, prob.all_lambda_coul, prob.all_lambda_vdw)
        ts=[]
        for _ in range(4):
            t=time.perf_counter(); ctx.set_list(prob.nblist, prob.nenergrp_pairs); ts.append(time.perf_counter()-t)
        print(name,"set_atoms %.2f ms"%(ta*1e3),"set_list ms",[round(x*1e3,2) for x in ts])
