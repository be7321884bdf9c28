#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_inference.py tests/test_gpu_optim.py -q > $OUT/pytest_c4.log 2>&1; echo "pytest rc=$?" )
( timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_dec_c4.log 2>&1; echo "time rc=$?" )
tail -4 $OUT/pytest_c4.log; grep -E "^iter" $OUT/time_dec_c4.log; grep -E "att-bwd" -A16 $OUT/time_dec_c4.log; grep -E "^att loop: cycles|^gen loop: cycles|gen-bwd" -A9 $OUT/time_dec_c4.log | head -50
