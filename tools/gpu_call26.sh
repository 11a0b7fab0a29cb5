#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call26.log
: > $L
echo "== pytest batch-related" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu -k "batch or many or random" 2>&1 | tail -3 >> $L
echo "== timeline" >> $L
timeout 600 python tools/prof_e2e_timeline.py > gpurun_out/s2_timeline2.log 2>&1; grep "=== call\|samples done\|sample 21 launched" gpurun_out/s2_timeline2.log >> $L
