#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call17.log
: > $L
echo "== pytest gpu (all)" >> $L
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 >> $L
echo "== bench default" >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2>> $L
echo "== bench reference arm" >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>> $L
echo "== ncu launch list of bench (skip e2e)" >> $L
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-300 >> $L
echo "== ncu full: fused sweep + traceback kernels of 256 resident samples" >> $L
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"dip_sweep4_many|dip_anc_many" -c 2 -o gpurun_out/r02_sweep4_many256 -f python tools/prof_v4_many.py 256 1 >> $L 2>&1
for f in gpurun_out/r02_bench_1gpu.json gpurun_out/r02_bench_reference_arm.json; do echo "$f: $(python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d.get('metric'), '%.4g'%d.get('value',0), 'ms', '%.1f'%d.get('ms_per_step',0), 'frac', d.get('roofline',{}).get('frac'), 'e2e', '%.4g'%d.get('e2e',{}).get('value',0), 'kernel_ms', d.get('kernel_ms'))
" 2>&1)" >> $L; done
