#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call21.log
: > $L
echo "== pytest diploid + sketch/hap quick" >> $L
timeout 1200 python -m pytest tests/test_dp_diploid_gpu.py tests/test_dp_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
echo "== bench default" >> $L
timeout 900 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/s2_bench_try.json 2>> $L
python -c "
import json
d=json.loads(open('gpurun_out/s2_bench_try.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms')}, d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_single_sample']['ms_per_step'])" >> $L 2>&1
echo "== c4 s4" >> $L
timeout 600 python tools/prof_c4.py 4 18 2>&1 | grep "^value" >> $L
