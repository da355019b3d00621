#!/bin/bash
# shared delay-line rows in the split / fused paths: parity on every path, then A/B against one row per speaker
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3
AW_FUSED_TILE=0 timeout 600 python -m pytest tests/test_gpu_convolution.py tests/test_gpu_eq.py -m gpu -x -q --timeout 600 2>&1 | tail -2
AW_PERSISTENT=0 timeout 600 python -m pytest tests/test_gpu_convolution.py tests/test_gpu_eq.py -m gpu -x -q --timeout 600 2>&1 | tail -2
b() { name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-100} --warmup 10 --no-cpu --e2e-steps 3 --no-single-block "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config'].get('blocks_per_step',1)
    print('$name'.ljust(20), 'value', round(d['value']), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), {a: round(b,4) for a,b in d['step_roofline']['kernels_ms'].items()})
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
b C5-4096_rows AW_X=0 -- --workload C5-4096
b C5-4096_sep AW_KP_MERGE_ROWS=0 -- --workload C5-4096
b C5-4096_rows2 AW_X=0 -- --workload C5-4096
b C2_split_rows AW_FUSED_TILE=0 -- --workload C2
b C2_split_sep AW_FUSED_TILE=0 AW_KP_MERGE_ROWS=0 -- --workload C2
b C2_fused_rows AW_PERSISTENT=0 -- --workload C2
b C2_fused_sep AW_PERSISTENT=0 AW_KP_MERGE_ROWS=0 -- --workload C2
