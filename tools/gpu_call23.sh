#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call23.log
: > $L
echo "== pytest gpu (all)" >> $L
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6 >> $L
echo "== bench default" >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2>> $L
echo "== bench reference arm" >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> $L
for R in 18 6 12 24 30 36; do
  echo "== bench c4_h90_s4 R$R" >> $L
  timeout 900 python bench.py --workload c4_h90_s4 --R $R --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/r02_bench_c4_h90_s4_R$R.json 2>> $L
done
for f in gpurun_out/r02_bench_1gpu.json gpurun_out/r02_bench_reference_arm.json gpurun_out/r02_bench_c4_h90_s4_R*.json; do echo "$f: $(python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d.get('metric'), '%.4g'%d.get('value',0), 'ms', '%.1f'%d.get('ms_per_step',0), 'frac', d.get('roofline',{}).get('frac'), 'traffic', d.get('roofline',{}).get('traffic'), 'e2e', '%.4g'%d.get('e2e',{}).get('value',0), 'kernel_ms', d.get('kernel_ms'))
" 2>&1)" >> $L; done
