#!/bin/bash
set -u
out=gpurun_out/r02_call36
mkdir -p $out
timeout 1200 python -m pytest -q -m gpu tests/test_mdrun_nb_gpu_route.py -k two_domain -p no:cacheprovider -x -rA > $out/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest.log | cut -c1-300; grep -E "^E  " $out/pytest.log | head -12 | cut -c1-400
# the cluster kernel at C5 size (1 M atoms), beside the reference's CUDA kernel
timeout 1500 python tools/nb_bench.py C5 --steps 10 --fork-gpu > $out/nb_C5.json 2> $out/nb_C5.err; echo "nb C5 rc=$?"; cut -c1-1500 $out/nb_C5.json; tail -2 $out/nb_C5.err
