#!/bin/bash
set -u
out=gpurun_out/r02_call31
mkdir -p $out
timeout 1500 python -m pytest -q -m gpu tests/test_mdrun_nb_gpu_route.py -k "steady and c3" -p no:cacheprovider -x -rA > $out/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest.log | cut -c1-200; grep -E "^E  " $out/pytest.log | head -10 | cut -c1-300
tail -6 gpurun_out/mdrun_nb_gpu_route_timing.txt | cut -c1-500
