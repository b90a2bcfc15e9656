#!/bin/bash
# Round 2 (2 GPUs): suite incl. the one-launch tile pass test; pair kernels after the ragged-chunk trim (+ ncu --set full
# of both); 2-GPU parity and cfg2 bench with the one-launch tile pass on / off.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02l_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02l_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02l_gpu_tests.log | cut -c1-300 | head -40
timeout 300 python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol > gpurun_out/r02l_mf_cfg2.json 2>&1
timeout 300 python scripts/matvec_free_bench.py > gpurun_out/r02l_mf_cfg5_slice.json 2>&1
tail -1 gpurun_out/r02l_mf_cfg2.json | cut -c1-330; tail -1 gpurun_out/r02l_mf_cfg5_slice.json | cut -c1-330
stamp "matrix-free operator"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02l_mg_check_n$N.log 2>&1; stamp "multi_gpu_check (peer) rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|Error|rror:|assert" gpurun_out/r02l_mg_check_n$N.log | head -20
for SM in 1 0; do
  timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --opt symop_multi=$SM > gpurun_out/r02l_bench_cfg2_n${N}_multi$SM.json 2> gpurun_out/r02l_bench_cfg2_n${N}_multi$SM.err; stamp "bench cfg2 n=$N symop_multi=$SM rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02l_bench_cfg2_n${N}_multi$SM.json')); print('n=$N multi=$SM value', d['value'], 'matvec ms', d['phases'].get('matvec_avg_ms'), d['phases']['per_step'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mv_pairs2_kernel -s 2 -c 1 -o gpurun_out/r02l_pairs2 -f \
    python scripts/matvec_free_bench.py --reps 1 > gpurun_out/r02l_ncu_pairs2.log 2>&1; stamp "ncu pairs2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mv_pairs_kernel -s 2 -c 1 -o gpurun_out/r02l_pairs1 -f \
    python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --reps 1 > gpurun_out/r02l_ncu_pairs1.log 2>&1; stamp "ncu pairs (cfg2) rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
