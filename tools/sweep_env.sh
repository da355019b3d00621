#!/bin/bash
# usage: tools/sweep_env.sh WORKLOAD "TAG1:VAR=VAL VAR2=VAL" "TAG2:..."   (run under gpurun)
WORKLOAD=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  env $envs python bench.py --workload $WORKLOAD --steps 100 --warmup 20 --no-cpu --e2e-steps 3 2>&1 | tail -1 > gpurun_out/sweep_${WORKLOAD}_$tag.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sweep_${WORKLOAD}_$tag.json'))
    print('$tag', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'step_frac', round(d['step_roofline']['frac'],3), d['config']['plan'], {k: round(v,4) for k,v in d['step_roofline']['kernels_ms'].items()})
except Exception as e:
    print('$tag', 'FAILED', e, open('gpurun_out/sweep_${WORKLOAD}_$tag.json').read()[-300:])
PY
done
