#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
b() { name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu --e2e-steps 3 "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config']['blocks_per_step']; sb=d.get('single_block_calls') or {}
    print('$name'.ljust(20), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), 'single', round(sb.get('ms_per_block',0),4), round(sb.get('frac',0),3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['step_roofline']['kernels_ms'])
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
b C4_overlap AW_X=0 -- --workload C4
b C4_serial AW_X=0 -- --workload C4 --serial-eq
b F3_overlap AW_X=0 -- --workload F3
b F3_serial AW_X=0 -- --workload F3 --serial-eq
b C2 AW_X=0 -- --workload C2
b C5-64 AW_X=0 -- --workload C5-64
b C5-512 AW_X=0 -- --workload C5-512
b C3 AW_X=0 -- --workload C3
b C4_overlap2 AW_X=0 -- --workload C4
