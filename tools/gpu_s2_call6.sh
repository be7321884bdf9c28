#!/bin/bash
# session-2 call 6: templated wide variants (no register regression on the common path), barrier-overlapped operand prefetch, singleton capture stream
set -u
OUT=gpurun_out; mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_s2c6.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_s2c6.log )
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
run() { tag=$1; shift; timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra-baselines "$@" > $OUT/bench_s2c6_$tag.log 2>&1; line $tag $OUT/bench_s2c6_$tag.log; }
run headline
run cfg1_ljspeech_B16 --config ljspeech --batch 16
timeout 300 python tools/time_decoder.py --B 60 --kind zoneout --precision bf16 --iters 2 > $OUT/time_decoder_s2c6.log 2>&1; grep -A28 "gen-bwd loop" $OUT/time_decoder_s2c6.log
