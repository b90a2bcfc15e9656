#!/bin/bash
# the driver's scaling command at N GPUs (cfg2, default mode), shortened to 1 warm-up + 2 steps
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r01o_bench_cfg2_n$N.json 2> gpurun_out/r01o_bench_cfg2_n$N.err
echo "bench cfg2 n=$N rc=$?"
cat gpurun_out/r01o_bench_cfg2_n$N.json | cut -c1-3000
grep -vE "^\*|OMP_NUM|^$" gpurun_out/r01o_bench_cfg2_n$N.err | tail -5
