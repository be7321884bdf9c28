#!/bin/bash
# session-2 call 2: persistent bi-LSTM + wider split-K: tests, cfg 1 / cfg 3 / cfg 4-5 bench lines, headline bench
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_bilstm.py tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_decoder.py -x -q > $OUT/pytest_s2c2.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_s2c2.log )
line() { python - "$1" "$2" <<'PY'
import sys, json
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    t = d.get('roofline', {}).get('timing', {})
    print(tag, 'ms/step', round(d['ms_per_step'], 2), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['workload'][:70],
          {k: round(v['ms_per_launch'], 2) for k, v in t.items()})
except Exception as e:
    print(tag, 'FAILED', e); print(open(path).read()[-1500:])
PY
}
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra-baselines "$@" > $OUT/bench_s2c2_$tag.log 2>&1; line $tag $OUT/bench_s2c2_$tag.log; }
run headline
run cfg3_shared_switching_B64 --config shared_switching --batch 64
B200TTS_BILSTM_CHAIN=1 run cfg3_chain --config shared_switching --batch 64
run cfg1_ljspeech_B16 --config ljspeech --batch 16
B200TTS_BILSTM_CHAIN=1 run cfg1_chain --config ljspeech --batch 16
run cfg45_generated_switching_B60_L300_T1200 --config generated_switching --batch 60 --text-len 300 --frames 1200
run cfg45_generated_switching_B80_L300_T1200 --config generated_switching --batch 80 --text-len 300 --frames 1200
run fp32_parity_mode --precision fp32 --steps 2 --warmup 1
