#!/bin/bash
export FEPB200_TIMING=1
c=${1:-C5}
python tools/time_compute.py $c 2>&1 | tail -3
FEPB200_HOST_THREADS=16 FEPB200_HOST_GRAIN=4096 python tools/time_compute.py $c 2>&1 | tail -3
FEPB200_HOST_THREADS=16 FEPB200_HOST_GRAIN=4096 FEPB200_ZC_OUT=1 python tools/time_compute.py $c 2>&1 | tail -3
