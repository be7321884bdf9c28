#!/bin/bash
# last verification of the final tree: full GPU suite, headline bench with baselines, other configs, kernel table, phase counters, loop traffic
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_final3.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_final3.log )
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_final3.log 2>&1; echo "bench rc=$?"; tail -1 $OUT/bench_final3.log | cut -c1-300
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-baselines "$@" > $OUT/bench_final3_$tag.log 2>&1; tail -1 $OUT/bench_final3_$tag.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['value']))"; }
run cfg3_shared_switching_B64 --config shared_switching --batch 64
run cfg1_ljspeech_B16 --config ljspeech --batch 16
run cfg45_generated_switching_B60_L300_T1200 --config generated_switching --batch 60 --text-len 300 --frames 1200
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extra-baselines --breakdown $OUT/breakdown_final3.txt > $OUT/bench_final3_bd.log 2>&1; echo "bd rc=$?"
timeout 300 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_decoder_final3.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --replay-mode application --clock-control none --metrics $M -k regex:'loop|att_post' --csv --log-file $OUT/ncu_loops_app_replay_final3.csv \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_loops_app_replay_final3.log 2>&1; echo "ncu loops rc=$?"
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python tools/time_decoder.py --B 8 --L 40 --T 6 --kind zoneout --precision bf16 --iters 0 > $OUT/sanitizer_final3_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -1 $OUT/sanitizer_final3_racecheck.log
