#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call4.log
: > $L
echo "== pytest diploid gpu" >> $L
timeout 900 python -m pytest tests/test_dp_diploid_gpu.py -x -q -m gpu 2>&1 | tail -5 >> $L
echo "== prof_v4 256" >> $L
timeout 600 python tools/prof_v4.py 256 2>&1 | grep -v "^\[" >> $L
echo "== cli x4" >> $L
timeout 900 python -m pytest tests/test_cli_gpu.py -x -q -m gpu -k "replicated" 2>&1 | tail -5 >> $L
