#!/bin/bash
set -u
out=gpurun_out/r02_call46
mkdir -p $out
timeout 200 python tests/parity_report.py > $out/parity_report_full_size.txt 2> $out/parity_report.err; echo "report rc=$?"; cat $out/parity_report_full_size.txt | cut -c1-900; tail -3 $out/parity_report.err | cut -c1-300
