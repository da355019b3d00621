#!/bin/bash
mkdir -p gpurun_out
b() { name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu --e2e-steps 3 "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config']['blocks_per_step']; sb=d.get('single_block_calls') or {}
    print('$name'.ljust(20), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), 'single', round(sb.get('ms_per_block',0),4), round(sb.get('frac',0),3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
L=$PWD/airwave_b200/lib
for w in C2 C5-64 C5-128; do
  b ${w}_base AW_X=0 -- --workload $w
  for v in P3 P2 P1; do b ${w}_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload $w; done
done
