#!/bin/bash
set -u
out=gpurun_out/r02_call47
mkdir -p $out
timeout 400 python -m pytest -q -m gpu tests/test_gpu_parity.py -k "full_size or push_reduction" -rA -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR)" $out/pytest.log | cut -c1-160; grep -E "^E  " $out/pytest.log | head -10 | cut -c1-300; tail -1 $out/pytest.log
