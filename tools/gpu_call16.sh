#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call16.log
: > $L
echo "== pytest diploid" >> $L
timeout 1200 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
echo "== many 256 (vup)" >> $L
timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 DG_NO_VUP" >> $L
DG_NO_VUP=1 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
