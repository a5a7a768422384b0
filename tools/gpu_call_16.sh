#!/bin/bash
set -u
out=gpurun_out/r02_call16
mkdir -p $out
rm -f gpurun_out/mdrun_gpu_route_timing.txt
timeout 1200 python -m pytest -q -m gpu tests/test_gpu_pairs14.py tests/test_mdrun_pairs14.py tests/test_mdrun_gpu_route.py tests/test_mdrun_gpu_build.py -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $out/pytest.log | cut -c1-300
cp gpurun_out/mdrun_gpu_route_timing.txt $out/ 2>/dev/null
grep -A4 "steady state" gpurun_out/mdrun_gpu_route_timing.txt | cut -c1-400
