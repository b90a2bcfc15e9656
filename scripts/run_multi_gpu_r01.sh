#!/bin/bash
# 2-GPU validation (run under gpurun --gpus 2): single-GPU parity tests of the new pieces, the sharded-vs-single
# parity check, and short sharded bench runs in both assembled modes.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r01_mg_parity_tests.log 2>&1; stamp "parity tests rc=$?"
tail -3 gpurun_out/r01_mg_parity_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r01_mg_check_n$N.log 2>&1; stamp "multi_gpu_check rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|Error|error|assert" gpurun_out/r01_mg_check_n$N.log | head -20
for mode in assembled assembled_sym; do
  timeout 600 $TR bench.py --gpus $N --steps 1 --warmup 0 --tol 1e-3 --mode $mode --no-cpu-baseline --e2e-steps 1 \
      > gpurun_out/r01_mg_bench_n${N}_$mode.json 2> gpurun_out/r01_mg_bench_n${N}_$mode.err; stamp "bench $mode rc=$?"
  tail -c 600 gpurun_out/r01_mg_bench_n${N}_$mode.json; echo
  tail -5 gpurun_out/r01_mg_bench_n${N}_$mode.err
done
