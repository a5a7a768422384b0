#!/bin/bash
set -u
out=gpurun_out/r02_call33
mkdir -p $out
for e in "" "--energy"; do
echo "== default $e"
timeout 600 python tools/nb_bench.py C3 --steps 30 $e 2>>$out/err.log | cut -c150-330
echo "== row skip $e"
FEPB200_LIB=$PWD/tools/ab/libfepb200_nb_rowskip.so timeout 600 python tools/nb_bench.py C3 --steps 30 $e 2>>$out/err.log | cut -c150-330
done
FEPB200_LIB=$PWD/tools/ab/libfepb200_nb_rowskip.so timeout 900 python -m pytest -q -m gpu tests/test_gpu_nb.py -p no:cacheprovider -x > $out/pytest_rowskip.log 2>&1; echo "pytest rowskip rc=$?"; tail -2 $out/pytest_rowskip.log
GMX_FEPB200_LIB_OVERRIDE=1 true
# in-route, c3_hexadecane, with the row-skip library
cd /tmp && mkdir -p r33 && cd r33 && export LD_LIBRARY_PATH=/root/repo/integration/_gmx_cuda/lib && for lib in /root/repo/gromacs-fep-gpu_b200/lib/libfepb200.so /root/repo/tools/ab/libfepb200_nb_rowskip.so; do GMX_ENABLE_GPU_TIMING=1 GMX_FEPB200_NB=1 GMX_FEPB200=1 GMX_FEPB200_LIB=$lib timeout 600 /root/repo/integration/_gmx_cuda/bin/gmx -quiet mdrun -s /root/repo/tests/golden/mdrun_tpr/c3_hexadecane.tpr -deffnm run -nb gpu -pme cpu -bonded cpu -update cpu -fep gpu -ntmpi 1 -ntomp 2 -notunepme -nsteps 600 -resethway > run.out 2>&1; echo "$lib: $(grep 'Nonbonded F' run.log | tr -s ' ')"; done
