#!/bin/bash
# f-3 first GPU run: parity tests of the cluster-pair kernel, a timing on C3, ncu --set full of the kernel
set -u
out=gpurun_out/r02_call21
mkdir -p $out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_nb.py -p no:cacheprovider -x > $out/pytest_nb.log 2>&1
echo "pytest nb rc=$?"; tail -15 $out/pytest_nb.log | cut -c1-400
timeout 600 python tools/nb_bench.py C3 --steps 20 > $out/nb_c3.json 2> $out/nb_c3.err; echo "nb bench rc=$?"; cat $out/nb_c3.json | cut -c1-900; tail -3 $out/nb_c3.err
timeout 600 python tools/nb_bench.py C3 --steps 20 --energy > $out/nb_c3_energy.json 2>> $out/nb_c3.err; echo "nb bench energy rc=$?"; cat $out/nb_c3_energy.json | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fep_nb_kernel -c 3 -o $out/prof_nb_c3 python tools/nb_bench.py C3 --steps 2 --warmup 1 > $out/ncu_nb.log 2>&1; echo "ncu rc=$?"
ls -la $out
