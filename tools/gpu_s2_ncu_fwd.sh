#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_loop_tc_kernel -c 2 -o $OUT/prof_fwd_loops_final \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 --fwd-only > $OUT/ncu_fwd_full_final.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT/prof_fwd_loops_final.ncu-rep
