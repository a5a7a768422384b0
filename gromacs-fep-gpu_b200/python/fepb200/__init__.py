"""fepb200 -- host-side Python layer of the B200 FEP perturbed-pair kernel.

Product modules: `lib` (ctypes binding of libfepb200.so, the C-ABI of include/fepb200.h),
`params`, `problem`, `shard`, `synth`.  Nothing in this package imports from `oracle/`.
"""
from .params import *  # noqa: F401,F403
from .problem import FepList, Problem, shift_vectors, nbfp_from_c6c12, nbfp_grid_geometric  # noqa: F401
