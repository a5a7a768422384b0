#!/bin/bash
# 2 GPUs: bench.py exactly as the driver launches it (default reduction = p2p-push), and the reference arm under torchrun
set -u
out=gpurun_out/r02_call48
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 100 --warmup 5 > $out/bench2.json 2> $out/bench2.err
echo "bench 2 rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_call48/bench2.json").read().strip().splitlines()[-1])
print(d["run"], "ms/step %.4f value %.3e e2e ms %.4f launches %d" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["gpu_launches"]), d["clocks"])
PY
