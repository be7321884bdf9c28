#!/bin/bash
# Round-2 evidence run (one gpurun call): GPU tests, compute-sanitizer on the persistent kernels at small T,
# ncu application-replay metrics of the cluster / cooperative loop kernels (kernel replay cannot restore them).
set -u
mkdir -p gpurun_out
OUT=gpurun_out
python -c "import torch; print(torch.cuda.get_device_name(0))" > $OUT/probe_env.txt 2>&1
nproc >> $OUT/probe_env.txt; lscpu | grep 'Model name' >> $OUT/probe_env.txt
( timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/probe_env.txt )
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/time_decoder.py --B 8 --L 40 --T 6 --kind zoneout --precision bf16 --iters 0 \
      > $OUT/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?" >> $OUT/probe_env.txt
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --replay-mode application --clock-control none --metrics $M -k regex:'loop|att_post' --csv --log-file $OUT/ncu_loops_app_replay.csv \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_loops_app_replay.log 2>&1
echo "ncu app replay rc=$?" >> $OUT/probe_env.txt
timeout 600 python bench.py --steps 5 --warmup 3 --breakdown $OUT/breakdown_r2_start.txt > $OUT/bench_r2_start.log 2>&1
echo "bench rc=$?" >> $OUT/probe_env.txt
cat $OUT/probe_env.txt
