#!/bin/bash
# session-2 call 16: Toeplitz pair build moved into the shadow of the query-partial loads
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_model.py tests/test_gpu_t900.py -x -q > $OUT/pytest_s2c16.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_s2c16.log )
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-baselines > $OUT/bench_s2c16.log 2>&1
tail -1 $OUT/bench_s2c16.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
timeout 300 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_decoder_s2c16.log 2>&1; grep -A22 "^att loop: cycles" $OUT/time_decoder_s2c16.log
