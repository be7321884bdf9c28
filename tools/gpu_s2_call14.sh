#!/bin/bash
# session-2 call 14: generator bottleneck products folded into the expand / tail kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_t900.py tests/test_gpu_modules.py tests/test_gpu_reference_train.py -x -q > $OUT/pytest_s2c14.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_s2c14.log )
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-baselines > $OUT/bench_s2c14.log 2>&1
tail -1 $OUT/bench_s2c14.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extra-baselines --breakdown $OUT/breakdown_s2c14.txt > $OUT/bench_s2c14_bd.log 2>&1; echo "bd rc=$?"
grep "bn_\|generator_\|TOTAL" $OUT/breakdown_s2c14.txt | head -20
