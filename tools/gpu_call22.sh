#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call22.log
: > $L
for R in 36 30; do
echo "== prof c4 s4 R$R" >> $L
timeout 600 python tools/prof_c4.py 4 $R 2>&1 | grep "^value" >> $L
done
echo "== pytest wide" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu -k "beyond or wide or full_size" 2>&1 | tail -2 >> $L
