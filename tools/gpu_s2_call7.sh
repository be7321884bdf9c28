#!/bin/bash
# session-2 call 7: zero-padded implicit convolutions (80-channel layers), pack_conv_weight indexing: conv / model tests, bench, breakdown
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_t900.py tests/test_gpu_optim.py -x -q > $OUT/pytest_s2c7.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_s2c7.log )
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-baselines > $OUT/bench_s2c7.log 2>&1
tail -1 $OUT/bench_s2c7.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extra-baselines --breakdown $OUT/breakdown_s2c7.txt > $OUT/bench_s2c7_bd.log 2>&1; echo "bd rc=$?"
