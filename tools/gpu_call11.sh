#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call11.log
: > $L
echo "== many 256 default" >> $L
timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 NCW=9" >> $L
DG_V4_NCW=9 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 RC=5" >> $L
DG_V4_RC=5 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 NSLOT=6 SLOT=4096" >> $L
DG_V4_NSLOT=6 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 SLOT=8192" >> $L
DG_V4_SLOT=8192 timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== pytest quick" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu -k "beyond or random or variants" 2>&1 | tail -3 >> $L
