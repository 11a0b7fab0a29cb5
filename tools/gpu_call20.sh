#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call20.log
: > $L
echo "== e2e batch tuning" >> $L
timeout 600 python tools/prof_e2e_batch.py 2>&1 | grep "workers" >> $L
echo "== c4 s4 NCW=16" >> $L
DG_V4_NCW=16 timeout 600 python tools/prof_c4.py 4 18 2>&1 | grep -v "^config\|^bench:" >> $L
echo "== c4 s4 NCW=8" >> $L
DG_V4_NCW=8 timeout 600 python tools/prof_c4.py 4 18 2>&1 | grep "^value" >> $L
