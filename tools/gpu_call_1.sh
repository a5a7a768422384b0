#!/bin/bash
# Round 2, first GPU call: every GPU test once (no -x: each file gets its verdict), the bench line,
# compute-sanitizer on smoke().  Everything is written under gpurun_out/r02_call1/.
set -u
out=gpurun_out/r02_call1
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest -q -m gpu tests -rA --durations=25 -p no:cacheprovider > $out/pytest_all.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_all.log
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest_all.log | sort | uniq -c | sort -rn | head -5
grep -E "^(FAILED|ERROR)" $out/pytest_all.log | head -40
tail -5 $out/pytest_all.log
timeout 600 python bench.py --steps 50 --warmup 5 > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; cat $out/bench.json | head -c 3000
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $out/sanitizer_memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -4 $out/sanitizer_memcheck_smoke.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > $out/sanitizer_racecheck_smoke.log 2>&1
echo "racecheck rc=$?"; tail -4 $out/sanitizer_racecheck_smoke.log
cp gpurun_out/*.txt $out/ 2>/dev/null
true
