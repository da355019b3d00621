#!/bin/bash
# K2 / K4: frame operands loaded without L1 allocation (twiddles stay in L1), K4 with two barriers less per frame
mkdir -p gpurun_out
L=$PWD/airwave_b200/lib
b() { name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-100} --warmup 10 --no-cpu --e2e-steps 3 --no-single-block "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config'].get('blocks_per_step',1)
    print('$name'.ljust(20), 'value', round(d['value']), 'ms/block', round(d['ms_per_step']/k,4), 'stepfrac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), {a: round(b,4) for a,b in d['step_roofline']['kernels_ms'].items()})
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 2>&1 | tail -1
AW_FUSED_TILE=0 timeout 600 python -m pytest tests/test_gpu_convolution.py tests/test_gpu_eq.py -m gpu -x -q --timeout 600 2>&1 | tail -1
for v in R3 N1 R3 N1; do b C5-4096_${v} AW_LIBRARY=$L/libairwave_$v.so -- --workload C5-4096; done
for v in R3 N1; do b C5-2048s_${v} AW_LIBRARY=$L/libairwave_$v.so AW_FUSED_TILE=0 -- --workload C5-2048; done
for v in R3 N1; do b C2s_${v} AW_LIBRARY=$L/libairwave_$v.so AW_FUSED_TILE=0 -- --workload C2; done
for v in R3 N1; do b C2f_${v} AW_LIBRARY=$L/libairwave_$v.so AW_PERSISTENT=0 -- --workload C2; done
