#!/bin/bash
# 2 GPUs: where the step's time goes on every rank (shard kernels / reduction), and the reduction kernel on its own
set -u
out=gpurun_out/r02_call39
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29504 tools/time_multi_gpu_phases.py C5 > $out/phases_p2p.txt 2> $out/phases_p2p.err
echo "phases rc=$?"; cat $out/phases_p2p.txt | cut -c1-700; tail -3 $out/phases_p2p.err | cut -c1-300
