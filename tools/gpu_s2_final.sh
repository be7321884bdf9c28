#!/bin/bash
# session-2 evidence run on the final revision: full GPU suite, headline bench (with CPU / eager baselines), other configs, kernel table,
# ncu: launch list of one step, DRAM traffic of the loops (application replay), tensor-pipe / DRAM of the largest GEMM / conv-block kernels,
# full-set capture of the attention reverse loop; example trainer
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import torch; print(torch.cuda.get_device_name(0))" > $OUT/final_env.txt 2>&1; nproc >> $OUT/final_env.txt
( timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_final.log )
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_final.log 2>&1; echo "bench rc=$?"; tail -1 $OUT/bench_final.log | cut -c1-400
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-baselines "$@" > $OUT/bench_final_$tag.log 2>&1; tail -1 $OUT/bench_final_$tag.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],2), round(d['value']), round(d['e2e']['value']))"; }
run cfg3_shared_switching_B64 --config shared_switching --batch 64
run cfg1_ljspeech_B16 --config ljspeech --batch 16
run cfg45_generated_switching_B60_L300_T1200 --config generated_switching --batch 60 --text-len 300 --frames 1200
timeout 600 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-extra-baselines --breakdown $OUT/breakdown_final.txt > $OUT/bench_final_bd.log 2>&1; echo "bd rc=$?"
timeout 300 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_decoder_final.log 2>&1
# ncu: every launch of one eager step with its device time (cold cache, serialised: shares only)
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file $OUT/launches_r2_final.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-extra-baselines > $OUT/launches_r2_final.log 2>&1; echo "ncu launches rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --replay-mode application --clock-control none --metrics $M -k regex:'loop|att_post' --csv --log-file $OUT/ncu_loops_app_replay_final.csv \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_loops_app_replay_final.log 2>&1; echo "ncu loops rc=$?"
timeout 900 ncu --clock-control none --metrics $M -k regex:'gemm_tc_kernel|block_fwd|block_bwd|bn_stats|bn_bwd|pack_im2col|pack_kcontig' -c 200 --csv --log-file $OUT/ncu_gemm_conv_final.csv \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-extra-baselines > $OUT/ncu_gemm_conv_final.log 2>&1; echo "ncu gemm rc=$?"
B200TTS_PROFILE_NO_COOP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_bwd_loop -c 1 -o $OUT/prof_att_bwd_final \
    python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 0 > $OUT/ncu_att_bwd_full_final.log 2>&1; echo "ncu full rc=$?"
timeout 600 python examples/train_synthetic.py --steps 6 > $OUT/train_synthetic_final.log 2>&1; echo "example rc=$?"; tail -3 $OUT/train_synthetic_final.log
