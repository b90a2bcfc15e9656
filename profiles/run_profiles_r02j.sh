#!/bin/bash
# Round 2 (2 GPUs): suite + bench on GPU 0 with the CUDA-graph pivot chunks; 2-GPU parity; cfg2 on 2 GPUs; which
# preconditioner form suits the loose tolerance of cfg5 (aspirin-size, M = 5000 stand-in on 2 GPUs).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-2}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02j_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02j_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02j_gpu_tests.log | cut -c1-300 | head -30
for G in 1 0; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --opt pchol_graph=$G > gpurun_out/r02j_bench_graph$G.json 2> gpurun_out/r02j_bench_graph$G.err
  python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_graph$G.json')); print('pchol_graph=$G value', d['value'], [ (s['pchol_build_s'], s['preconditioner_s'], s['cg_iters']) for s in d['phases']['per_step']])"
done
stamp "graph A/B (1 GPU)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02j_mg_check_n$N.log 2>&1; stamp "multi_gpu_check (peer) rc=$?"
grep -E "MULTI_GPU_CHECK|iters sharded|pivots|Woodbury|projected form|Error|rror:|assert" gpurun_out/r02j_mg_check_n$N.log | head -30
MLFFPC_PEER=0 NCCL_DEBUG=WARN timeout 600 $TR tests/multi_gpu_check.py > gpurun_out/r02j_mg_check_nccl_n$N.log 2>&1; stamp "multi_gpu_check (nccl) rc=$?"
grep -E "MULTI_GPU_CHECK|Error|rror:|assert" gpurun_out/r02j_mg_check_nccl_n$N.log | head -10
timeout 600 $TR bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt > gpurun_out/r02j_bench_cfg2_n$N.json 2> gpurun_out/r02j_bench_cfg2_n$N.err; stamp "bench cfg2 n=$N rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_cfg2_n$N.json')); print('n=$N value', d['value'], d['phases']['per_step'])"
for F in projected woodbury; do
  timeout 900 $TR bench.py --gpus $N --workload cfg5 --M 5000 --k 4096 --tol 1e-4 --mode matrix_free --precon-form $F --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-alt \
     > gpurun_out/r02j_cfg5m5000_$F.json 2> gpurun_out/r02j_cfg5m5000_$F.err; stamp "cfg5-like $F rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02j_cfg5m5000_$F.json')); p=d['phases']; print('$F value', d['value'], 'iters', p['cg_iters'], 'precon', p['preconditioner_s'], 'pchol', p['pchol_build_s'], 'cg', p['cg_s'], 'apply ms', p['precon_apply_avg_ms'], 'conv', p['converged'])"
  grep -vE "^\*|OMP_NUM|^$" gpurun_out/r02j_cfg5m5000_$F.err | tail -3
done
