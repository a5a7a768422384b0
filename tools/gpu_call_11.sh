#!/bin/bash
# 2 GPUs: the multi-GPU tests (fused, p2p = reduce-scatter, p2p-allreduce, nccl) and the bench line per reduction.
set -u
out=gpurun_out/r02_call11
mkdir -p $out
timeout 600 python -m pytest -q -m gpu tests/test_multi_gpu.py tests/test_gpu_parity.py tests/test_gpu_peer_exchange.py -x -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $out/pytest.log | cut -c1-300
for red in p2p p2p-allreduce fused; do
  FEPB200_REDUCTION=$red FEPB200_E2E_PHASES=1 timeout 300 python bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/bench2_$red.json 2> $out/bench2_$red.err
  echo "bench $red rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/bench2_$red.json").read().strip().splitlines()[-1])
    print("$red", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), d["kernel_ms"], d["run"]["reduction"])
except Exception as e: print("parse failed", e)
PY
  grep "e2e phases" $out/bench2_$red.err | head -2
done
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-fork-gpu > $out/bench1.json 2> $out/bench1.err; echo "bench1 rc=$?"; cut -c1-1500 $out/bench1.json
