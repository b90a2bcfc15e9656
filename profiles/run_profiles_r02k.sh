#!/bin/bash
# Round 2 (1 GPU): suite after the narrow contraction tiles / whole-wave split-K, the left-looking TRSM and the Gram fold
# interval; A/B of each on cfg2; per-kernel times of the matrix-free operator (ncu launch list).
set -u
mkdir -p gpurun_out
T0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_gpu_tests.log 2>&1; stamp "pytest -m gpu rc=$?"
grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02k_gpu_tests.log | tail -30
grep -E "^E  " gpurun_out/r02k_gpu_tests.log | cut -c1-300 | head -40
for NV in 1 0; do
  MLFFPC_DGEMM_NARROW=$NV timeout 300 python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol > gpurun_out/r02k_mf_cfg2_narrow$NV.json 2>&1
  MLFFPC_DGEMM_NARROW=$NV timeout 300 python scripts/matvec_free_bench.py > gpurun_out/r02k_mf_cfg5_slice_narrow$NV.json 2>&1
  tail -1 gpurun_out/r02k_mf_cfg2_narrow$NV.json | cut -c1-330; tail -1 gpurun_out/r02k_mf_cfg5_slice_narrow$NV.json | cut -c1-330
done
stamp "contraction tiles A/B"
MLFFPC_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02k_bench_default.json 2> gpurun_out/r02k_bench_default.err; stamp "bench default rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --opt gram_fold=1 > gpurun_out/r02k_bench_fold1.json 2> gpurun_out/r02k_bench_fold1.err; stamp "bench gram_fold=1 rc=$?"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-alt --opt gram_fold=4 > gpurun_out/r02k_bench_fold4.json 2> gpurun_out/r02k_bench_fold4.err; stamp "bench gram_fold=4 rc=$?"
python - <<'PY'
import json
for tag in ('default', 'fold1', 'fold4'):
    try:
        d = json.load(open('gpurun_out/r02k_bench_%s.json' % tag))
        print(tag, 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'alt', d.get('alt'))
        for s in d['phases']['per_step']: print('  ', s)
    except Exception as e:
        print(tag, 'parse failed', e)
PY
grep -E "gram|trsm|potrf|Mk|defect|factor" gpurun_out/r02k_bench_default.err | tail -24
# per-kernel times of the matrix-free operator (serialised, cold cache: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02k_mf_cfg5_launches.csv \
    python scripts/matvec_free_bench.py --reps 2 > gpurun_out/r02k_mf_ncu5.log 2>&1; stamp "ncu launch list cfg5 slice rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02k_mf_cfg2_launches.csv \
    python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --reps 2 > gpurun_out/r02k_mf_ncu2.log 2>&1; stamp "ncu launch list cfg2 rc=$?"
python - <<'PY'
import csv, io
for tag in ('cfg5', 'cfg2'):
    try:
        lines = open('gpurun_out/r02k_mf_%s_launches.csv' % tag).read().splitlines()
        i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
        rows = [r for r in csv.DictReader(io.StringIO('\n'.join(lines[i:]))) if r['Metric Name'] == 'gpu__time_duration.sum']
        print(tag, 'last 12 launches:')
        for r in rows[-12:]:
            print('   %-60s grid %-16s %8.1f us' % (r['Kernel Name'][:60], r['Grid Size'].replace(' ', ''), float(r['Metric Value']) / 1e3))
    except Exception as e:
        print(tag, 'launch list parse failed', e)
PY
