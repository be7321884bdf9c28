#!/bin/bash
# 2-GPU sanity of the driver's launch line on the final revision
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_final_2gpu.log 2>&1; echo "rc=$?"; tail -1 $OUT/bench_final_2gpu.log | cut -c1-500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > $OUT/bench_final_2gpu_ref.log 2>&1; echo "ref rc=$?"; tail -1 $OUT/bench_final_2gpu_ref.log | cut -c1-300
