#!/bin/bash
mkdir -p gpurun_out
b() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu --e2e-steps 3 --no-single-block "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config']['blocks_per_step']
    print('$name'.ljust(28), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
L=$PWD/airwave_b200/lib
for w in C5-1024 C5-2048; do
  b ${w}_base AW_X=0 -- --workload $w
  for v in F384P2 F384P2PF F512P2 F640P2; do
    b ${w}_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload $w
  done
  b ${w}_base2 AW_X=0 -- --workload $w
done
# correctness of the candidate geometries
for v in F512P2 F640P2 F384P2PF; do
  AW_LIBRARY=$L/libairwave_$v.so timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 300 -k "1024 or 2048" 2>&1 | tail -2
done
