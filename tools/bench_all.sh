#!/bin/bash
# GPU-side helper: one bench line per workload (run under gpurun); extra env via $BENCH_ENV, extra flags via $BENCH_ARGS
mkdir -p gpurun_out
for w in "$@"; do
  env $BENCH_ENV python bench.py --workload $w --steps ${STEPS:-400} --warmup 40 --no-cpu --e2e-steps 5 $BENCH_ARGS 2>gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$w.json'))
    k=d['config']['blocks_per_step']; sb=d.get('single_block_calls') or {}
    print('$w', 'value', round(d['value']), 'blocks/call', k, 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3),
          '| one block per call: value', round(sb.get('value',0)), 'ms/block', round(sb.get('ms_per_block',0),4), 'frac', round(sb.get('frac',0),3), 'p99', round((sb.get('latency_ms') or {}).get('p99',0),4),
          '| e2e', round(d['e2e']['value']), d['config']['plan'], {a: round(b,4) for a,b in d['step_roofline']['kernels_ms'].items()}, d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$w', 'FAILED', e, open('gpurun_out/bench_$w.err').read()[-300:])
PY
done
