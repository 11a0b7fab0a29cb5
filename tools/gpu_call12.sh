#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call12.log
: > $L
for v in nogiant_pairs nogiant_hoist giant_pairs giant_hoist; do
  echo "== many 256 $v" >> $L
  DG_CUDA_LIB_OVERRIDE=$PWD/tools/ab/lib_$v.so timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
done
echo "== many 256 nogiant_pairs again" >> $L
DG_CUDA_LIB_OVERRIDE=$PWD/tools/ab/lib_nogiant_pairs.so timeout 300 python tools/prof_v4_many.py 256 3 2>&1 | tail -1 >> $L
