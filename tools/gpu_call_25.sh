#!/bin/bash
set -u
out=gpurun_out/r02_call25
mkdir -p $out
timeout 1200 python -m pytest -q -m gpu tests/test_gpu_nb.py -p no:cacheprovider -x > $out/pytest_nb.log 2>&1
echo "pytest rc=$?"; tail -5 $out/pytest_nb.log | cut -c1-400
for e in "" "--energy"; do
echo "== v3 (i atoms in registers, L1 prefetch, 128 regs) $e"
FEPB200_LIB=$PWD/tools/ab/libfepb200_nb_v3.so timeout 600 python tools/nb_bench.py C3 --steps 30 $e 2>>$out/err.log | cut -c150-330
echo "== v4 96 regs, 5 CTAs/SM $e"
timeout 600 python tools/nb_bench.py C3 --steps 30 $e 2>>$out/err.log | cut -c150-330
echo "== v4 122 regs, 4 CTAs/SM $e"
FEPB200_NB_CTAS_PER_SM=4 timeout 600 python tools/nb_bench.py C3 --steps 30 $e 2>>$out/err.log | cut -c150-330
done
echo "== C2, C4, C5 with the default"
for c in C2 C4; do timeout 900 python tools/nb_bench.py $c --steps 20 2>>$out/err.log | cut -c1-420; done
