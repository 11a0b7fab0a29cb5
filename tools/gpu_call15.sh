#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call15.log
: > $L
echo "== config 5 (backbone 625 kbp, 46-walk graphs, 22 samples)" >> $L
timeout 1500 python tests/perf/run_config5.py --backbone 625000 --sites 4125 --ref-samples 2 > gpurun_out/r02_config5_625k.json 2>> $L
tail -c 1500 gpurun_out/r02_config5_625k.json >> $L
for R in 18 6 12 24 30 36; do
  echo "== bench c4_h90_s4 R$R" >> $L
  timeout 900 python bench.py --workload c4_h90_s4 --R $R --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/r02_bench_c4_h90_s4_R$R.json 2>> $L
  python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_c4_h90_s4_R$R.json').read().strip().splitlines()[-1])
print('R$R', '%.4g'%d['value'], 'ms %.1f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))" >> $L 2>&1
done
