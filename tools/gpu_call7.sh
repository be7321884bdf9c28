#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_t900.py -q -x -k "bf16 or t900" > $OUT/pytest_c7.log 2>&1; echo "pytest rc=$?" )
( timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_dec_c7.log 2>&1; echo "time rc=$?" )
tail -3 $OUT/pytest_c7.log; grep -E "^iter" $OUT/time_dec_c7.log; grep -E "att-bwd" -A16 $OUT/time_dec_c7.log | head -17
