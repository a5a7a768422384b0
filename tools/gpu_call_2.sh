#!/bin/bash
# new trip layout: correctness first (parity + list tests), then timing
set -u
out=gpurun_out/r02_call2
mkdir -p $out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_parity.py -x -rA -p no:cacheprovider > $out/pytest_parity.log 2>&1
echo "parity rc=$?"; grep -E "^(FAILED|ERROR)" $out/pytest_parity.log | head -20; tail -15 $out/pytest_parity.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_call2/bench.json').readline())
    print('step us', d['ms_per_step']*1e3, 'e2e us', d['e2e']['ms_per_step']*1e3, d['roofline']['kernel_ms'])
except Exception as e:
    print('no bench line', e); print(open('gpurun_out/r02_call2/bench.err').read()[-2000:])
PY
