#!/bin/bash
# cfg5 matrix-free operator: per-rank slice timing + ncu launch list (one GPU).
set -u
mkdir -p gpurun_out
C="python scripts/matvec_free_bench.py --reps 3"
timeout 300 $C > gpurun_out/r01i_mf_cfg5_slice.json 2> gpurun_out/r01i_mf.err && {
  cat gpurun_out/r01i_mf_cfg5_slice.json
  timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'mv_|dgemm' -c 20 --csv \
      --log-file gpurun_out/r01i_mf_launches.csv $C > gpurun_out/r01i_mf_ncu_stdout.log 2>&1
  echo "ncu rc=$?"
  grep -v "^==" gpurun_out/r01i_mf_launches.csv | python -c "
import csv,sys
for x in csv.DictReader(sys.stdin): print(x['Kernel Name'][:60], x['Grid Size'], x['Metric Value'], x['Metric Unit'])"
}
tail -3 gpurun_out/r01i_mf.err
python scripts/matvec_free_bench.py --M 4000 --world 1 --kind ethanol --reps 3
