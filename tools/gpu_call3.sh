#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_loss.py tests/test_gpu_modules.py tests/test_gpu_inference.py tests/test_gpu_reference_train.py tests/test_gpu_optim.py tests/test_gpu_decoder.py tests/test_gpu_t900.py -q -k "not gemm and not attention_step" > $OUT/pytest_c3.log 2>&1; echo "pytest rc=$?" )
( timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_dec_c3.log 2>&1; echo "time rc=$?" )
( B200TTS_ATT_BWD_MMA_SYNC=1 timeout 600 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 1 > $OUT/time_dec_c3_mmasync.log 2>&1; echo "time(mma.sync) rc=$?" )
( timeout 600 python tools/time_decoder.py --B 60 --L 300 --T 1200 --kind zoneout --precision bf16 --iters 1 > $OUT/time_dec_c3_L300.log 2>&1; echo "time(L300) rc=$?" )
tail -5 $OUT/pytest_c3.log; grep -E "^iter|att-bwd" -A8 $OUT/time_dec_c3.log | head -40
