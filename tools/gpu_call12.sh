#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_c12.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_c12.log )
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --breakdown $OUT/breakdown_c12.txt > $OUT/bench_c12.log 2>&1
tail -1 $OUT/bench_c12.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print(round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config']['launch'][:40], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
sed -n 6,40p $OUT/breakdown_c12.txt
