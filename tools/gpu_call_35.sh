#!/bin/bash
# 2-GPU sanity of what the driver runs at round end for N > 1: the multi-GPU tests and bench.py under torchrun
set -u
out=gpurun_out/r02_call35
mkdir -p $out
timeout 600 python -m pytest -q -m gpu tests/test_multi_gpu.py -p no:cacheprovider > $out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -2 $out/pytest_multi.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 100 --warmup 5 > $out/bench2.json 2> $out/bench2.err; echo "bench 2 rc=$?"
python - <<PY
import json
d=json.loads(open("$out/bench2.json").read().strip().splitlines()[-1])
print("N=2", d["run"]["reduction"], "ms/step %.4f value %.3e e2e ms %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), d.get("cluster_pair_kernel"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29503 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $out/bench2_ref.json 2> $out/bench2_ref.err; echo "bench ref 2 rc=$?"; cut -c1-300 $out/bench2_ref.json
