#!/bin/bash
# What the first GPU call of the next round should run (everything below was written after round 1's GPU budget
# was spent; each step writes its notes under gpurun_out/):
#   gpurun --timeout 2400 -- 'bash tools/round2_first_gpu_call.sh'
set -u
mkdir -p gpurun_out
# 1. the tests that have not run on a B200 yet, in the order of tests/ (no -x: every file gets its verdict)
python -m pytest -q -m gpu tests/test_z1_gpu_device_handoff.py tests/test_z2_mdrun_dropin_more.py tests/test_z3_fork_cuda.py \
  tests/test_z4_mdrun_gpu_build.py tests/test_z5_mdrun_gpu_route.py tests/test_z6_mdrun_pairs14.py -rA > gpurun_out/round2_pending_tests.log 2>&1
tail -40 gpurun_out/round2_pending_tests.log
# 2. the bench line with the fork's own CUDA kernels beside ours (fork_gpu_baseline) and the full-size comparison
python bench.py --steps 50 --warmup 5 > gpurun_out/round2_bench.json 2> gpurun_out/round2_bench.err
python tests/fork_cuda_compare.py C5 C2 C4g1 > gpurun_out/round2_fork_cuda_compare.jsonl 2>&1
# 3. the A/B runs of the switches prepared in round 1
bash tools/experiments_round2.sh C5 > gpurun_out/round2_experiments.log 2>&1
tail -60 gpurun_out/round2_experiments.log
