#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in tc mmasync nores tc2; do
  case $v in
    tc|tc2) envs="";;
    mmasync) envs="B200TTS_ATT_BWD_MMA_SYNC=1";;
    nores) envs="B200TTS_ATT_BWD_NO_RESIDENT=1";;
  esac
  env $envs timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c9_$v.log 2>&1
  echo "$v rc=$?"
  tail -1 $OUT/bench_c9_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('$v', round(d['ms_per_step'],2), {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
done
