#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call19.log
: > $L
nvidia-smi --query-gpu=memory.total,memory.used --format=csv >> $L
echo "== pytest gpu (all)" >> $L
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 >> $L
echo "== ncu launch list of the resident step (bench --skip-e2e)" >> $L
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/r02_ncu_bench.log 2>&1
tail -1 gpurun_out/r02_ncu_bench.log | cut -c1-200 >> $L
