#!/bin/bash
set -u
out=gpurun_out/r02_call17
mkdir -p $out
FEPB200_TIMING=1 timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-fork-gpu --no-side-configs > $out/bench_timing.json 2> $out/bench_timing.err
echo "bench timing rc=$?"; grep -v "set_list" $out/bench_timing.err | tail -12 | cut -c1-400
timeout 600 python bench.py --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
print("value %.3e ms %.4f e2e %.4f ms" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]))
print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("kernels","note","peak_source")})
for k in d["roofline"]["kernels"]: print("  ", {a:b for a,b in k.items() if a!="ncu"})
print("every_step", d["every_step"])
print("cpu", d["cpu_baseline"])
print("fork", d["fork_gpu_baseline"])
for n,c in (d.get("configs") or {}).items(): print(n, {a:b for a,b in c.items() if a!="config"})
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 $out/bench_ref.json
