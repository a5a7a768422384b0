#!/bin/bash
set -u
out=gpurun_out/r02_call30
mkdir -p $out
timeout 1500 python -m pytest -q -m gpu tests/test_gpu_nb.py tests/test_mdrun_nb_gpu_route.py -p no:cacheprovider -x -rA > $out/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest.log | cut -d' ' -f1 | sort | uniq -c; grep -E "^(FAILED|ERROR)" $out/pytest.log | cut -c1-250; grep -E "^E  " $out/pytest.log | head -20 | cut -c1-300
cat gpurun_out/direct_forces_timing.txt 2>/dev/null

tail -6 gpurun_out/mdrun_nb_gpu_route_timing.txt | cut -c1-500
