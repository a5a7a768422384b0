#!/bin/bash
set -u
out=gpurun_out/r02_call23
mkdir -p $out
timeout 1200 python -m pytest -q -m gpu tests/test_gpu_nb.py tests/test_gpu_device_handoff.py -p no:cacheprovider -x > $out/pytest_nb.log 2>&1
echo "pytest rc=$?"; tail -5 $out/pytest_nb.log | cut -c1-400
for c in C3 C2; do
timeout 600 python tools/nb_bench.py $c --steps 20 > $out/nb_$c.json 2> $out/nb_$c.err; echo "nb bench $c rc=$?"; cat $out/nb_$c.json | cut -c1-1200; tail -3 $out/nb_$c.err
done
timeout 600 python tools/nb_bench.py C3 --steps 20 --energy > $out/nb_C3_energy.json 2>> $out/nb_C3.err; echo "nb bench energy rc=$?"; cat $out/nb_C3_energy.json | cut -c1-1200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fep_nb_kernel -c 3 -o $out/prof_nb_c3 python tools/nb_bench.py C3 --steps 2 --warmup 1 > $out/ncu_nb.log 2>&1; echo "ncu rc=$?"
