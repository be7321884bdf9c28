#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python bench.py --steps 8 --warmup 3 --no-extra-baselines --breakdown $OUT/breakdown_c8.txt > $OUT/bench_c8.log 2>&1; echo "bench rc=$?" )
tail -1 $OUT/bench_c8.log | cut -c1-400; head -30 $OUT/breakdown_c8.txt
