#!/bin/bash
# session-2 call 4: in-place gradient accumulation, shared dw fragment table, fast tanh: full GPU suite, bench, breakdown, ncu DRAM traffic of the loops
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_s2c4.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_s2c4.log )
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-baselines > $OUT/bench_s2c4.log 2>&1
tail -1 $OUT/bench_s2c4.log | python -c "
import sys,json
d=json.loads(sys.stdin.read())
t=d['roofline']['timing']
print('ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], {k:round(v['ms_per_launch'],2) for k,v in t.items()})"
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extra-baselines --breakdown $OUT/breakdown_s2c4.txt > $OUT/bench_s2c4_bd.log 2>&1; echo "bd rc=$?"
timeout 300 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_decoder_s2c4.log 2>&1; grep -A8 "att-bwd loop" $OUT/time_decoder_s2c4.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --replay-mode application --clock-control none --metrics $M -k regex:'loop|att_post' --csv --log-file $OUT/ncu_loops_app_replay_s2c4.csv \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_loops_app_replay_s2c4.log 2>&1
echo "ncu rc=$?"
