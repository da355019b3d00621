#!/bin/bash
# KP at B >= 1024: fewer producer warps, the registers to the FFT / MAC warps (KA: 2 producers, 112 registers; KB: + operand
# prefetch; KC: 3 producers, 104; KD: 1 producer, 120) against the shipped 4 producers, 96 registers
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
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
for v in KA KD; do AW_LIBRARY=$L/libairwave_$v.so timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 -k "1024 or 2048" 2>&1 | tail -1; done
for w in C5-1024 C5-2048; do for v in cuda KA KB KC KD cuda KA KD; do b ${w}_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload $w; done; done
