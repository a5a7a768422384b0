#!/bin/bash
set -u
out=gpurun_out/r02_call24
mkdir -p $out
timeout 1200 python -m pytest -q -m gpu tests/test_gpu_nb.py -p no:cacheprovider -x > $out/pytest_nb.log 2>&1
echo "pytest rc=$?"; tail -5 $out/pytest_nb.log | cut -c1-400
timeout 600 python tools/nb_bench.py C3 --steps 20 --cpu-baseline --fork-gpu > $out/nb_C3.json 2> $out/nb_C3.err; echo "nb bench C3 rc=$?"; cat $out/nb_C3.json | cut -c1-2000; tail -3 $out/nb_C3.err
FEPB200_NB_CTAS_PER_SM=4 timeout 600 python tools/nb_bench.py C3 --steps 20 > $out/nb_C3_occ4.json 2>> $out/nb_C3.err; echo "occ4 rc=$?"; cat $out/nb_C3_occ4.json | cut -c1-500
timeout 600 python tools/nb_bench.py C2 --steps 20 > $out/nb_C2.json 2> $out/nb_C2.err; echo "nb bench C2 rc=$?"; cat $out/nb_C2.json | cut -c1-500
timeout 600 python tools/nb_bench.py C3 --steps 20 --energy --fork-gpu > $out/nb_C3_energy.json 2>> $out/nb_C3.err; echo "nb bench energy rc=$?"; cat $out/nb_C3_energy.json | cut -c1-2000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fep_nb_kernel -c 3 -o $out/prof_nb_c3 python tools/nb_bench.py C3 --steps 2 --warmup 1 > $out/ncu_nb.log 2>&1; echo "ncu rc=$?"
