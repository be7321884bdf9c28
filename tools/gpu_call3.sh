#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_gpu_decoder.py -x -q -k "loss or bf16_perf or golden" > $OUT/pytest_c3.log 2>&1; echo "pytest rc=$?" )
( timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_dec_c3.log 2>&1; echo "time rc=$?" )
( B200TTS_ATT_BWD_MMA_SYNC=1 timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 1 > $OUT/time_dec_c3_mmasync.log 2>&1; echo "time(mma.sync) rc=$?" )
tail -30 $OUT/time_dec_c3.log
