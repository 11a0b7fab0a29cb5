#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call14.log
: > $L
echo "== many 256 lpc1" >> $L
DG_CUDA_LIB_OVERRIDE=$PWD/tools/ab/lib_lpc1.so timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== many 256 product (lpc2)" >> $L
timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
echo "== pytest diploid + cli c4" >> $L
timeout 1200 python -m pytest tests/test_dp_diploid_gpu.py tests/test_cli_gpu.py -x -q -m gpu -k "not mhc_hg002 and not vcf and not 18" 2>&1 | tail -4 >> $L
echo "== prof_v4 (single + packed alone) product" >> $L
timeout 600 python tools/prof_v4.py 2>&1 | grep -v "^\[" >> $L
echo "== prof_v4 lpc1" >> $L
DG_CUDA_LIB_OVERRIDE=$PWD/tools/ab/lib_lpc1.so timeout 600 python tools/prof_v4.py 2>&1 | grep -v "^\[" >> $L
echo "== prof c4 s4" >> $L
timeout 600 python tools/prof_c4.py 4 18 2>&1 | grep -v "^config\|^bench:" >> $L
