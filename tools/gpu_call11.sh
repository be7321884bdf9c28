#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_modules.py -q -x > $OUT/pytest_c11.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_c11.log )
for v in graph nograph graph2; do
  extra=""; [ $v = nograph ] && extra="--no-graph"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $extra > $OUT/bench_c11_$v.log 2>&1
  tail -1 $OUT/bench_c11_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('$v', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config']['launch'][:60], d['gpu_launches'], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
done
