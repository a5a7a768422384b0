#!/bin/bash
# the driver's round-end sequence, rehearsed on 8 GPUs: multi-GPU tests, bench at N = 1, 2, 4, 8 (default reduction), reference arm
set -u
out=gpurun_out/r02_call20
mkdir -p $out
timeout 600 python -m pytest -q -m gpu tests/test_multi_gpu.py -p no:cacheprovider > $out/pytest_multi.log 2>&1
echo "pytest multi rc=$?"; tail -3 $out/pytest_multi.log | cut -c1-300
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-fork-gpu --no-side-configs > $out/bench$n.json 2> $out/bench$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 100 --warmup 5 > $out/bench$n.json 2> $out/bench$n.err
  fi
  echo "bench $n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out/bench$n.json").read().strip().splitlines()[-1])
    print("N=$n", d["run"]["reduction"], "ms/step %.4f value %.3e e2e ms %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]), {k:round(v*1e3,1) for k,v in d["kernel_ms"].items()}, "force-only %.4f" % d["every_step"]["ms_per_step"], "roofline frac", d["roofline"]["frac"], d["roofline"]["frac_is"][:40])
except Exception as e: print("parse failed", e)
PY
done
FEPB200_REDUCTION=fused timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 > $out/bench8_fused.json 2> $out/bench8_fused.err; echo "fused rc=$?"
FEPB200_REDUCTION=p2p-allreduce timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 > $out/bench8_allreduce.json 2> $out/bench8_allreduce.err; echo "allreduce rc=$?"
python - <<PY
import json
for f in ("bench8_fused","bench8_allreduce"):
    try:
        d=json.loads(open("$out/%s.json" % f).read().strip().splitlines()[-1]); print(f, "ms/step %.4f e2e %.4f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
    except Exception as e: print(f, "parse failed", e)
PY
