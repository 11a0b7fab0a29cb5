#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call6.log
: > $L
for w in c4_h90_s16 c4_h90_s4; do
  echo "== bench $w R18" >> $L
  timeout 1500 python bench.py --workload $w --R 18 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/s2_${w}_R18.json 2>> $L
  tail -c 3000 gpurun_out/s2_${w}_R18.json >> $L
  nvidia-smi --query-gpu=memory.used --format=csv >> $L
done
