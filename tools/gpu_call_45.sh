#!/bin/bash
# 4 GPUs: the two-rank check of tests/test_multi_gpu.py with four ranks (push and pull), then bench.py at N = 4 with either
set -u
out=gpurun_out/r02_call45
mkdir -p $out
timeout 300 python tools/check_multi_gpu_world.py 4 p2p-push p2p > $out/check4.txt 2> $out/check4.err; echo "check rc=$?"; cat $out/check4.txt; grep -iE "error|Traceback|fault" $out/check4.err | head -5
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
  FEPB200_REDUCTION=$red timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 4 --steps 100 --warmup 5 > $out/bench4_$red.json 2> $out/bench4_$red.err
  echo "bench 4 $red rc=$?"; show $out/bench4_$red.json
done
