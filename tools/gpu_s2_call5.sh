#!/bin/bash
# session-2 call 5: memory dim 512 on the persistent kernels, graph captured on the warm-up stream: full GPU suite, headline + cfg 1 bench, sanitizer on the bi-LSTM
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_s2c5.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_s2c5.log )
line() { python - "$1" "$2" <<'PY'
import sys, json
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    t = d.get('roofline', {}).get('timing', {})
    print(tag, 'ms/step', round(d['ms_per_step'], 2), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'traffic', d['roofline'].get('traffic'),
          {k: round(v['ms_per_launch'], 2) for k, v in t.items()})
except Exception as e:
    print(tag, 'FAILED', e); print(open(path).read()[-1500:])
PY
}
run() { tag=$1; shift; timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-baselines "$@" > $OUT/bench_s2c5_$tag.log 2>&1; line $tag $OUT/bench_s2c5_$tag.log; }
run headline
run cfg1_ljspeech_B16 --config ljspeech --batch 16
run cfg1_ljspeech_B52 --config ljspeech --batch 52
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/bilstm_small.py > $OUT/sanitizer_bilstm_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -2 $OUT/sanitizer_bilstm_$tool.log
done
