#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
B200TTS_PROFILE_NO_COOP=1 timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SpeedOfLight --clock-control none --import-source on \
    -k regex:att_bwd_loop -c 1 -o $OUT/prof_att_bwd_r2 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_att_bwd.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu_att_bwd.log
