#!/bin/bash
# round-2 session-2 call 1: per-class cycle counters of the v4 sweep, and the A/B of what makes the fused sweep slower inside bench.py
mkdir -p gpurun_out
{
echo "== prof_v4 144 256"; timeout 600 python tools/prof_v4.py 144 256 2>&1 | grep -v "^\[" 
for o in "" "maxconn" "torch" "torch flush" "torch flush maxconn"; do
  echo "== many 256 3 $o"; timeout 300 python tools/prof_v4_many.py 256 3 $o 2>&1 | tail -1
done
} > gpurun_out/s2_call1.log 2>&1
