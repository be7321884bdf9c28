#!/bin/bash
# session-2 call 11: bf16 gate-gradient histories written by the reverse loops, read in place by the dX (K-major) and dW (MN-major) products
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_t900.py tests/test_gpu_optim.py -x -q > $OUT/pytest_s2c11.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_s2c11.log )
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-baselines > $OUT/bench_s2c11.log 2>&1
tail -1 $OUT/bench_s2c11.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extra-baselines --breakdown $OUT/breakdown_s2c11.txt > $OUT/bench_s2c11_bd.log 2>&1; echo "bd rc=$?"
grep "pack_\|TOTAL" $OUT/breakdown_s2c11.txt | head -20
