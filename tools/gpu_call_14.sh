#!/bin/bash
set -u
out=gpurun_out/r02_call14
mkdir -p $out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_peer_exchange.py tests/test_gpu_pairs14.py tests/test_gpu_device_handoff.py tests/test_host_cpp.py tests/test_mdrun_dropin.py tests/test_mdrun_dropin_more.py -x -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 $out/pytest.log | cut -c1-300
for w in all nofor force; do python tools/prof_step.py C5 20 $w; done 2>&1 | tee $out/times.txt
python tools/prof_step.py C2 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C4 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C3 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C3 20 all x nf=20 2>&1 | tee -a $out/times.txt
FEPB200_GAPSYS_GENERIC=1 python tools/prof_step.py C3 20 all x nf=20 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C1 20 all 2>&1 | tee -a $out/times.txt
FEPB200_TIMING=1 python tools/time_set_list.py 2>&1 | tail -12 | tee $out/set_list.txt
