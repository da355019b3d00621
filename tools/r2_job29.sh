#!/bin/bash
# KP at B = 1024 with 16 warps (5 transforms side by side, 128 registers per thread): V1 = 2 producers, V2 = + operand prefetch,
# V4 = 1 producer + prefetch; against the shipped 20 warps x 96 registers
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
    print('$name'.ljust(20), 'value', round(d['value']), 'ms/block', round(d['ms_per_step']/k,4), 'stepfrac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-300:])
PY
}
for v in V1 V2; do AW_LIBRARY=$L/libairwave_$v.so timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 -k "1024" 2>&1 | tail -1; done
for v in cuda V1 V2 V4 cuda V1 V2 V4; do b C5-1024_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload C5-1024; done
for v in cuda V1; do b C5-1024k4_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload C5-1024 --blocks-per-call 4; done
