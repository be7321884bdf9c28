#!/bin/bash
# session-2 baseline: GPU tests, bench (graphed + breakdown), reference arm
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_s2c1.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_s2c1.log )
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_s2c1.log 2>&1; echo "bench rc=$?"; tail -1 $OUT/bench_s2c1.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --breakdown $OUT/breakdown_s2c1.txt > $OUT/bench_s2c1_bd.log 2>&1; echo "bench bd rc=$?"; tail -1 $OUT/bench_s2c1_bd.log | cut -c1-600
head -40 $OUT/breakdown_s2c1.txt
