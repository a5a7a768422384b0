#!/bin/bash
set -u
out=gpurun_out/r02_call37
mkdir -p $out
timeout 1800 python -m pytest -q -m gpu tests/test_gpu_nb.py tests/test_mdrun_nb_gpu_route.py tests/test_mdrun_nb.py -p no:cacheprovider -rA > $out/pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(PASSED|FAILED|ERROR|SKIPPED)" $out/pytest.log | cut -d' ' -f1 | sort | uniq -c; grep -E "^(FAILED|ERROR)" $out/pytest.log | cut -c1-250; grep -E "^E  " $out/pytest.log | head -20 | cut -c1-300
timeout 600 python tools/nb_bench.py C3 --steps 20 > $out/nb_C3.json 2>$out/nb.err; cut -c150-330 $out/nb_C3.json
