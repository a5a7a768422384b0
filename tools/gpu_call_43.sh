#!/bin/bash
# 1 GPU: the cluster kernel's deviations with and without the adversarial placements; every GPU test, no -x
set -u
out=gpurun_out/r02_call43
mkdir -p $out
timeout 600 python tools/nb_parity_report.py > $out/nb_parity_report.txt 2> $out/nb_parity_report.err; echo "report rc=$?"; cat $out/nb_parity_report.txt | cut -c1-700; tail -3 $out/nb_parity_report.err | cut -c1-300
timeout 1500 python -m pytest -q -m gpu tests -rA --durations=10 -p no:cacheprovider > $out/pytest_all.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_all.log
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest_all.log | cut -d' ' -f1 | sort | uniq -c
grep -E "^(FAILED|ERROR)" $out/pytest_all.log | head -20 | cut -c1-250; grep -E "^E  " $out/pytest_all.log | head -30 | cut -c1-300
