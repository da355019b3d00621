#!/bin/bash
# GPU-side helper: one bench line per workload (run under gpurun); extra env via $BENCH_ENV
mkdir -p gpurun_out
for w in "$@"; do
  env $BENCH_ENV python bench.py --workload $w --steps 100 --warmup 40 --no-cpu --e2e-steps 5 2>&1 | tail -1 > gpurun_out/bench_$w.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    print('$w', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'step_frac', round(d['step_roofline']['frac'],3), 'e2e', round(d['e2e']['value']), d['config']['plan'], {k: round(v,4) for k,v in d['step_roofline']['kernels_ms'].items()}, 'p99', round(d['latency_ms']['p99'],4))
except Exception as e:
    print('$w', 'FAILED', e, open('gpurun_out/bench_$w.json').read()[-300:])
PY
done
