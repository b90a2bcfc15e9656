#!/bin/bash
# Round 2 (8 GPUs): what moves the CG iteration count of cfg2 on 8 GPUs (979 in r02h, 1123-1140 in r02n/o, 920 on one
# GPU)?  A/B: right-looking TRSM order of round 1, exact defect matrix E.
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() {  # tag, env, extra args
  env $2 timeout 300 $TR bench.py --gpus $N --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --north-star off $3 > gpurun_out/r02p_$1.json 2> gpurun_out/r02p_$1.err; stamp "$1 rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02p_$1.json')); print('$1 value', d['value'], d['phases']['per_step'], ' '.join('%.1e'%v for v in d['phases']['rel_resid_every_100_iters'][7:]))"
}
run trsm_right MLFFPC_TRSM_RIGHT=1 ""
run defect_exact MLFFPC_X=0 "--opt defect_mode=2"
