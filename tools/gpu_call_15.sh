#!/bin/bash
# everything once: all GPU tests (no -x: every file gets its verdict), incl. the mdrun drop-in / GPU-route runs with the rebuilt binaries
set -u
out=gpurun_out/r02_call15
mkdir -p $out
timeout 1500 python -m pytest -q -m gpu tests -rA --durations=15 -p no:cacheprovider > $out/pytest_all.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_all.log
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest_all.log | cut -d' ' -f1 | sort | uniq -c
grep -E "^(FAILED|ERROR)" $out/pytest_all.log | head -30 | cut -c1-250
tail -3 $out/pytest_all.log
cp gpurun_out/mdrun_*_timing.txt $out/ 2>/dev/null
true
