#!/bin/bash
# 2 GPUs: the light cross-GPU barrier (default) against the original one (FEPB200_BARRIER=sc), push and pull reductions
set -u
out=gpurun_out/r02_call41
mkdir -p $out
timeout 900 python -m pytest -q -m gpu tests/test_multi_gpu.py tests/test_gpu_parity.py -k "two_ranks or push_reduction or peer_reduce" -p no:cacheprovider -rA > $out/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest.log | cut -c1-200; grep -E "^E  " $out/pytest.log | head -20 | cut -c1-300; tail -2 $out/pytest.log
for bar in light sc; do for red in p2p-push p2p; do
  FEPB200_BARRIER=$bar FEPB200_REDUCTION=$red timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29504 tools/time_multi_gpu_phases.py C5 > $out/phases_${red}_$bar.txt 2> $out/phases_${red}_$bar.err
  echo "phases $red barrier=$bar rc=$?"; grep -E "^rank|^max" $out/phases_${red}_$bar.txt | cut -c1-600; grep -iE "error|fault|Traceback" $out/phases_${red}_$bar.err | head -5
done; done
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["run"]["reduction"], "ms/step %.4f (aligned %s) value %.3e e2e ms %.4f" % (d["ms_per_step"], d["run"].get("ms_per_step_aligned"), d["value"], d["e2e"]["ms_per_step"]),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()}, "force-only %.4f" % d["every_step"]["ms_per_step"])
except Exception as e:
    print("parse failed", sys.argv[1], e)
PY
}
for red in p2p-push p2p; do
  FEPB200_REDUCTION=$red timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 100 --warmup 5 > $out/bench2_$red.json 2> $out/bench2_$red.err
  echo "bench 2 $red rc=$?"; show $out/bench2_$red.json
done
