#!/bin/bash
set -u
out=gpurun_out/r02_call44
mkdir -p $out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_nb.py -rA -p no:cacheprovider > $out/pytest_nb.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_nb.log
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest_nb.log | cut -d' ' -f1 | sort | uniq -c
grep -E "^(FAILED|ERROR)" $out/pytest_nb.log | head -20 | cut -c1-250; grep -E "^E  " $out/pytest_nb.log | head -30 | cut -c1-300
