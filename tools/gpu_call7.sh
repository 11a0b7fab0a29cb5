#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call7.log
: > $L
echo "== pytest diploid gpu" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu 2>&1 | tail -5 >> $L
echo "== many 256 stride 512" >> $L
timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 stride 680" >> $L
DG_V4_STRIDE=680 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 stride 1024" >> $L
DG_V4_STRIDE=1024 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== bench c4_h90_s16 R18" >> $L
timeout 1500 python bench.py --workload c4_h90_s16 --R 18 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>> $L | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','kernel_ms','dp_value','recombinations')}, d['roofline']['frac'], d['e2e']['ms_per_step'])" >> $L 2>&1
