#!/bin/bash
# cfg4 (BASELINE.json configs[3]): n = 270 000 assembled in symmetric tile storage, sharded over N GPUs.
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 1200 $TR bench.py --gpus $N --steps 1 --warmup 0 --workload cfg4 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r01h_bench_cfg4_n$N.json 2> gpurun_out/r01h_bench_cfg4_n$N.err
echo "bench cfg4 n=$N rc=$?"
tail -c 2500 gpurun_out/r01h_bench_cfg4_n$N.json; echo
grep -vE "^\*|OMP_NUM|^$" gpurun_out/r01h_bench_cfg4_n$N.err | tail -8
