#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call2.log
: > $L
run() { echo "== $*" >> $L; timeout 400 "$@" 2>>$L | python -c "
import sys, json
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.rstrip()); continue
    print({k: d[k] for k in ('value','ms_per_step','kernel_ms')}, d['roofline']['frac'], d['config']['samples_per_gpu'], d['config']['distinct_samples'])
" >> $L 2>&1; }
run python bench.py --skip-e2e --distinct 1 --steps 2 --no-cpu-baseline
run python bench.py --skip-e2e --distinct 8 --steps 2 --no-cpu-baseline
run python bench.py --distinct 1 --steps 2 --no-cpu-baseline --e2e-steps 1
run python tools/prof_v4_many.py 296 3
run python tools/prof_v4_many.py 288 3
