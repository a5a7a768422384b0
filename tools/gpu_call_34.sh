#!/bin/bash
# final evidence, 1 GPU: every GPU test, the bench line, the ncu launch list of the bench command, ncu --set full of the kernels
set -u
out=gpurun_out/r02_call34
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest -q -m gpu tests -rA --durations=10 -p no:cacheprovider > $out/pytest_all.log 2>&1
echo "pytest rc=$?" | tee -a $out/pytest_all.log
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest_all.log | cut -d' ' -f1 | sort | uniq -c
grep -E "^(FAILED|ERROR)" $out/pytest_all.log | head -20 | cut -c1-250
timeout 600 python bench.py --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() { name=$1; skip=$2; cnt=$3; shift 3
  python tools/prof_step.py "$@" > $out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"fep_beutler_kernel|fep_epilogue|fep_gapsys|fep_pass_kernel|fep_foreign_kernel" -s $skip -c $cnt -o $out/prof_$name python tools/prof_step.py "$@" > $out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"; }
cap c5 9 3 C5 3 all
cap c4 6 2 C4 3 all
cap c2 6 2 C2 3 all
cap c3 9 3 C3 3 all x nf=20
ls -la $out/*.ncu-rep
# the cluster-pair kernel
timeout 600 python tools/nb_bench.py C3 --steps 20 --cpu-baseline --fork-gpu > $out/nb_C3.json 2> $out/nb_C3.err; echo "nb bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fep_nb_kernel -c 2 -o $out/prof_nb_c3 python tools/nb_bench.py C3 --steps 1 --warmup 1 > $out/ncu_nb.log 2>&1; echo "ncu nb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fep_nb_kernel -c 2 -o $out/prof_nb_c3_energy python tools/nb_bench.py C3 --steps 1 --warmup 1 --energy > $out/ncu_nb_e.log 2>&1; echo "ncu nb energy rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke.log
# compute-sanitizer: closed on this pool in earlier calls of the round; one more attempt on the smoke test
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 $out/sanitizer_memcheck_smoke.log | cut -c1-200
