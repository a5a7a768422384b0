#!/bin/bash
# A/B runs prepared at the end of round 1 (no GPU time was left to measure them).  Every knob defaults
# to the measured configuration; each line prints step / e2e / per-kernel device times of bench.py.
#   tools/experiments_round2.sh [C5]
c=${1:-C5}
run() { echo "== $*"; env "$@" tools/quick_bench.sh $c; }
run FEPB200_BASELINE=1
# 1. lanes per light atom in the epilogue (heavy atoms have their own role now): fewer blocks / waves
run FEPB200_EPI_LANES=4
run FEPB200_EPI_LANES=2
# 2. pass and foreign grids co-resident on every SM instead of one full wave after the other
#    (pass: 85 registers x 128 threads, foreign: 167 x 128; 2 + 2 CTAs need <= 64k registers)
run FEPB200_PASS_CTAS_PER_SM=3 FEPB200_FOREIGN_CTAS_PER_SM=1
run FEPB200_PASS_CTAS_PER_SM=2 FEPB200_FOREIGN_CTAS_PER_SM=2
run FEPB200_PASS_CTAS_PER_SM=2 FEPB200_FOREIGN_CTAS_PER_SM=1
# 3. pass + all foreign points from one load of each pair also on large lists
run FEPB200_FUSE=1
run FEPB200_FUSE=1 FEPB200_FOREIGN_CTAS_PER_SM=1
# 4. the two together
run FEPB200_EPI_LANES=4 FEPB200_PASS_CTAS_PER_SM=2 FEPB200_FOREIGN_CTAS_PER_SM=2
