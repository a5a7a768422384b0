#!/bin/bash
set -u
out=gpurun_out/r02_call3
mkdir -p $out
for w in all nofor force; do python tools/prof_step.py C5 20 $w; python tools/prof_step.py C5 20 $w flush; done 2>&1 | tee $out/times.txt
FEPB200_STAGE=direct python tools/prof_step.py C5 20 nofor 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C2 20 all 2>&1 | tee -a $out/times.txt
python tools/prof_step.py C5 3 nofor > $out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fep_beutler_kernel|fep_epilogue" -s 6 -c 4 -o $out/prof_c5_nofor python tools/prof_step.py C5 3 nofor > $out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/ncu.log
ls -la $out
