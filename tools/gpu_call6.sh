#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_c6.log 2>&1; echo "pytest rc=$?" )
( timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_dec_c6.log 2>&1; echo "time rc=$?" )
( B200TTS_ATT_BWD_NO_RESIDENT=1 timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 1 > $OUT/time_dec_c6_nores.log 2>&1; echo "time(no resident) rc=$?" )
tail -4 $OUT/pytest_c6.log; grep -E "^iter" $OUT/time_dec_c6.log; grep -E "att-bwd" -A16 $OUT/time_dec_c6.log | head -17; grep -E "^iter" $OUT/time_dec_c6_nores.log
