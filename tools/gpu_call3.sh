#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call3.log
: > $L
echo "== pytest diploid gpu" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu 2>&1 | tail -5 >> $L
echo "== prof_v4 256" >> $L
timeout 600 python tools/prof_v4.py 256 2>&1 | grep -v "^\[" >> $L
echo "== ncu set full (many 296)" >> $L
timeout 900 ncu --set full --import-source on --clock-control none -k regex:dip_sweep4_many -c 1 -o gpurun_out/s2_many296 -f python tools/prof_v4_many.py 296 1 >> $L 2>&1
