#!/bin/bash
# N-GPU validation (run under gpurun --gpus N): the sharded-vs-single parity check (small system), and sharded
# bench runs of the headline workload (tol 1e-6) in the default mode.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
WL=${2:-cfg2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r01s_mg_check_n$N.log 2>&1; stamp "multi_gpu_check rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|pivots|Woodbury|Error|rror:|assert" gpurun_out/r01s_mg_check_n$N.log | head -20
timeout 900 $TR bench.py --gpus $N --steps 2 --warmup 1 --workload $WL --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/r01s_mg_bench_${WL}_n$N.json 2> gpurun_out/r01s_mg_bench_${WL}_n$N.err; stamp "bench $WL n=$N rc=$?"
tail -c 1500 gpurun_out/r01s_mg_bench_${WL}_n$N.json; echo
grep -vE "^\*|OMP_NUM|^$" gpurun_out/r01s_mg_bench_${WL}_n$N.err | tail -5
