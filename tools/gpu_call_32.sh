#!/bin/bash
# ncu of the cluster kernels INSIDE the fork's mdrun -nb gpu (c3_hexadecane): ours and the fork's own, on the fork's pruned list
set -u
out=$PWD/gpurun_out/r02_call32
mkdir -p $out
GMX=$PWD/integration/_gmx_cuda/bin/gmx
export LD_LIBRARY_PATH=$PWD/integration/_gmx_cuda/lib:${LD_LIBRARY_PATH:-}
export GMX_FEPB200_LIB=$PWD/gromacs-fep-gpu_b200/lib/libfepb200.so
TPR=$PWD/tests/golden/mdrun_tpr/c3_hexadecane.tpr
ARGS="-quiet mdrun -s $TPR -deffnm run -nb gpu -pme cpu -bonded cpu -update cpu -fep gpu -ntmpi 1 -ntomp 2 -notunepme -nsteps 60"
mkdir -p /tmp/r32a /tmp/r32b
cd /tmp/r32a && GMX_FEPB200_NB=1 GMX_FEPB200=1 timeout 600 ncu --set full --clock-control none -k regex:fep_nb_kernel -s 40 -c 2 -o $out/prof_inroute_ours $GMX $ARGS > $out/ncu_ours.log 2>&1; echo "ncu ours rc=$?"
cd /tmp/r32b && timeout 600 ncu --set full --clock-control none -k regex:nbnxn_kernel_Elec -s 40 -c 2 -o $out/prof_inroute_fork $GMX $ARGS > $out/ncu_fork.log 2>&1; echo "ncu fork rc=$?"
ls -la $out
