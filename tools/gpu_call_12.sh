#!/bin/bash
# 8 GPUs: bench line for p2p (reduce-scatter) and fused at N = 8 and p2p at N = 4; the nccl 2-rank test again.
set -u
out=gpurun_out/r02_call12
mkdir -p $out
timeout 300 python -m pytest -q -m gpu tests/test_multi_gpu.py -x -p no:cacheprovider -k "nccl or p2p" > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $out/pytest.log | cut -c1-300
run() { n=$1; red=$2
  FEPB200_REDUCTION=$red FEPB200_E2E_PHASES=1 timeout 300 python bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/bench${n}_$red.json 2> $out/bench${n}_$red.err
  echo "bench $n $red rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/bench${n}_$red.json").read().strip().splitlines()[-1])
    print("$n $red", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "force-only", round(d["every_step"]["ms_per_step"],4))
except Exception as e: print("parse failed", e)
PY
  grep "e2e phases" $out/bench${n}_$red.err | head -1
}
run 8 p2p
run 8 fused
run 4 p2p
