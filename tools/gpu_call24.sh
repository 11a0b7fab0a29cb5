#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call24.log
: > $L
echo "== pytest gpu (all but the 160-GB case)" >> $L
timeout 2400 python -m pytest tests -q -m gpu -k "not stress-18 and not 18]" 2>&1 | tail -4 >> $L
echo "== bench default" >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2>> $L
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_single_sample']['ms_per_step'], d['cpu_baseline']['value'])" >> $L 2>&1
