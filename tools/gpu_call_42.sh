#!/bin/bash
# final evidence of round 2, 1 GPU (what the driver runs at round end, plus the profiles): every GPU test, smoke(), the bench line,
# the ncu launch list of the bench command, ncu --set full of the FEP kernels (the sources changed: PushTargets in KernelArgs)
set -u
out=gpurun_out/r02_call42
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest -q -m gpu tests -x -rA --durations=10 -p no:cacheprovider > $out/pytest_all.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_all.log
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest_all.log | cut -d' ' -f1 | sort | uniq -c
grep -E "^(FAILED|ERROR)" $out/pytest_all.log | head -20 | cut -c1-250; grep -E "^E  " $out/pytest_all.log | head -20 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.log | cut -c1-300
cap() { name=$1; skip=$2; cnt=$3; shift 3
  python tools/prof_step.py "$@" > $out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"fep_beutler_kernel|fep_epilogue|fep_gapsys|fep_pass_kernel|fep_foreign_kernel" -s $skip -c $cnt -o $out/prof_$name python tools/prof_step.py "$@" > $out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"; }
cap c5 9 3 C5 3 all
cap c4 6 2 C4 3 all
cap c2 6 2 C2 3 all
cap c3 9 3 C3 3 all x nf=20
ls -la $out/*.ncu-rep
# the counters bench.py quotes, tied to the sources that are running (the file is rebuilt here as well, from the same reports)
python tools/ncu_counters.py C5=$out/prof_c5.ncu-rep C4=$out/prof_c4.ncu-rep C2=$out/prof_c2.ncu-rep C3=$out/prof_c3.ncu-rep --note "gpurun call 42 of round 2; C3 with 20 optional foreign lambda" > $out/ncu_counters.log 2>&1; echo "ncu_counters rc=$?"
cp profiles/r02_ncu_counters.json profiles/r02_ncu_full_c?_raw.csv $out/
timeout 600 python bench.py --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference arm rc=$?"; cut -c1-300 $out/bench_reference.json
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
