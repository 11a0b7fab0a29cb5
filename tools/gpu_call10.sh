#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call10.log
: > $L
echo "== pytest diploid gpu + cli c4/x4" >> $L
timeout 1200 python -m pytest tests/test_dp_diploid_gpu.py tests/test_cli_gpu.py -x -q -m gpu -k "not mhc_hg002 and not vcf" 2>&1 | tail -4 >> $L
echo "== many 256" >> $L
timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== prof c4 s16" >> $L
timeout 600 python tools/prof_c4.py 16 18 2>&1 | grep -v "^config\|^bench:" >> $L
echo "== prof c4 s4" >> $L
timeout 600 python tools/prof_c4.py 4 18 2>&1 | grep -v "^config\|^bench:" >> $L
