#!/bin/bash
set -u
out=gpurun_out/r02_call18
mkdir -p $out
for zc in 1 0; do
  FEPB200_ZC_OUT=$zc FEPB200_TIMING=1 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/bench_zc$zc.json 2> $out/bench_zc$zc.err
  python - <<PY
import json
d=json.loads(open("$out/bench_zc$zc.json").read().strip().splitlines()[-1])
print("ZC_OUT=$zc: e2e %.4f ms, device %.4f ms" % (d["e2e"]["ms_per_step"], d["ms_per_step"]))
PY
  grep "compute()" $out/bench_zc$zc.err | cut -c1-300
done
for zc in 1 0; do for c in C2 C4; do
  FEPB200_ZC_OUT=$zc python tools/time_e2e_phases.py $c C5 2>&1 | tail -2 | cut -c1-300
done; done
