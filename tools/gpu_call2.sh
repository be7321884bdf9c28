#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_t900.py -x -q -s > $OUT/pytest_t900.log 2>&1; echo "t900 rc=$?" )
( timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_r2_b.log 2>&1; echo "bench rc=$?" )
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_r2_ref.log 2>&1; echo "ref rc=$?" )
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --replay-mode application --clock-control none --metrics $M -k regex:'loop|att_post' --csv --log-file $OUT/ncu_loops_app_replay.csv \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_loops_app_replay.log 2>&1
echo "ncu app replay (no coop) rc=$?"
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'att_bwd_loop|lstm_loop_tc' -c 2 -o $OUT/prof_att_loops_r2 \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_full_att.log 2>&1
echo "ncu full (no coop) rc=$?"
