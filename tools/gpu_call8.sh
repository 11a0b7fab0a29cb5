#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call8.log
: > $L
echo "== pytest diploid gpu" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
for st in 512 680 1024; do
echo "== many 256 stride $st" >> $L
DG_V4_STRIDE=$st timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
done
echo "== prof c4 s16" >> $L
timeout 600 python tools/prof_c4.py 16 18 2>&1 | grep -v "^config\|^bench:" >> $L
