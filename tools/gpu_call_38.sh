#!/bin/bash
# 2 GPUs: the multi-GPU tests, then bench.py at N = 2 with the ranks aligned on the device before every timed step
# (run.rank_alignment; the line carries the unaligned time too), default reduction and the fused exchange
set -u
out=gpurun_out/r02_call38
mkdir -p $out
timeout 600 python -m pytest -q -m gpu tests/test_multi_gpu.py -p no:cacheprovider > $out/pytest_multi.log 2>&1
echo "pytest multi rc=$?"; tail -3 $out/pytest_multi.log | cut -c1-300
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["run"]["reduction"], "ms/step %.4f (unaligned %s) value %.3e e2e ms %.4f" % (d["ms_per_step"], d["run"].get("ms_per_step_unaligned"), d["value"], d["e2e"]["ms_per_step"]),
          {k: round(v * 1e3, 1) for k, v in d["kernel_ms"].items()}, "force-only %.4f" % d["every_step"]["ms_per_step"])
except Exception as e:
    print("parse failed", sys.argv[1], e)
PY
}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 100 --warmup 5 > $out/bench2.json 2> $out/bench2.err
echo "bench 2 rc=$?"; show $out/bench2.json; tail -3 $out/bench2.err | cut -c1-300
FEPB200_REDUCTION=fused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --gpus 2 --steps 100 --warmup 5 > $out/bench2_fused.json 2> $out/bench2_fused.err
echo "bench 2 fused rc=$?"; show $out/bench2_fused.json
