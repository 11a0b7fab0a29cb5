#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/s2_call18.log
: > $L
nvidia-smi -L >> $L
echo "== smoke" >> $L; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> $L 2>&1
echo "== bench --gpus 2" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2>> $L
tail -c 600 gpurun_out/r02_bench_2gpu.json | cut -c1-600 >> $L
echo "== bench --mode row-sharded (2 GPUs)" >> $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode row-sharded --steps 3 --warmup 2 > gpurun_out/r02_bench_row_sharded_2gpu.json 2>> $L
tail -c 1500 gpurun_out/r02_bench_row_sharded_2gpu.json >> $L
echo "== pytest sharded (two processes, two GPUs)" >> $L
timeout 900 python -m pytest tests/test_dp_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
