#!/bin/bash
set -u
out=gpurun_out/r02_call10
mkdir -p $out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_peer_exchange.py tests/test_gpu_pairs14.py tests/test_gpu_device_handoff.py tests/test_host_cpp.py -x -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $out/pytest.log | cut -c1-300
for w in all nofor force; do python tools/prof_step.py C5 20 $w; done 2>&1 | tee $out/times.txt
FEPB200_STAGE=direct python tools/prof_step.py C5 20 nofor 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C2 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C4 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C3 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C5 3 nofor > $out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fep_beutler_kernel|fep_epilogue" -s 6 -c 2 -o $out/prof_c5_nofor python tools/prof_step.py C5 3 nofor > $out/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 $out/ncu.log
