#!/bin/bash
# Diagnose the 8-ranks-on-one-device failure of the fused exchange (trip layout v2): same test under four settings.
set -u
out=gpurun_out/r02_call6
mkdir -p $out
T="tests/test_gpu_peer_exchange.py"
run() { name=$1; shift; env "$@" timeout 300 python -m pytest -q -m gpu $T -p no:cacheprovider > $out/$name.log 2>&1; echo "$name rc=$?"; grep -E "passed|failed" $out/$name.log | tail -1; grep -E "fepb200 error" $out/$name.log | head -3; }
run default A=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run direct FEPB200_STAGE=direct
run nopdl FEPB200_OVERLAP=none
