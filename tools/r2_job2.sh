#!/bin/bash
# GPU-side (under gpurun): full GPU test-suite + same-box A/B of round-1 (in _r1_baseline/) vs the working tree
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
b() { # name, dir, env..., -- args
  name=$1; dir=$2; shift; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (cd $dir && env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-400} --warmup 40 --no-cpu --e2e-steps 3 "$@" 2>$OLDPWD/gpurun_out/b_$name.err | tail -1 > $OLDPWD/gpurun_out/b_$name.json)
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    print('$name', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['step_roofline']['frac'],3), 'p99', round(d['latency_ms']['p99'],4), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
for w in C2 C5-64 C5-128 C5-512 C5-1024 C5-2048 C3 C4; do
  b ${w}_old _r1_baseline AW_X=0 -- --workload $w
  b ${w}_new . AW_X=0 -- --workload $w
done
STEPS=100
for w in C2 C5-512; do
  b ${w}_old_burst _r1_baseline AW_X=0 -- --workload $w
  b ${w}_new_burst . AW_X=0 -- --workload $w
done
