#!/bin/bash
# Full-size (cfg2: n = 108 000, k = 4839) sharded-vs-single parity on N GPUs.
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
MG_M=4000 MG_K=4839 NCCL_DEBUG=WARN timeout 900 $TR tests/multi_gpu_check.py > gpurun_out/r01_mg_check_full_n$N.log 2>&1
echo "multi_gpu_check full rc=$?"
grep -vE "^\*|OMP_NUM_THREADS|^$" gpurun_out/r01_mg_check_full_n$N.log | tail -40
