#!/bin/bash
# Does a uniform shared-memory carve-out cure the 8-ranks-on-one-device hang?  And what does it cost / buy on C5?
set -u
out=gpurun_out/r02_call7
mkdir -p $out
T="tests/test_gpu_peer_exchange.py"
run() { name=$1; shift; env "$@" timeout 300 python -m pytest -q -m gpu $T -p no:cacheprovider > $out/$name.log 2>&1; echo "$name rc=$?"; grep -E "passed|failed" $out/$name.log | tail -1; grep -E "fepb200 error" $out/$name.log | head -2; }
run carve100 A=1
run carve_driver FEPB200_CARVEOUT=-1
run carve50 FEPB200_CARVEOUT=50
for cv in 100 -1 50; do
  for w in all nofor force; do FEPB200_CARVEOUT=$cv python tools/prof_step.py C5 20 $w; done
  FEPB200_CARVEOUT=$cv python tools/prof_step.py C2 20 all
done 2>&1 | tee $out/times.txt
