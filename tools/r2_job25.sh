#!/bin/bash
# K4 with the inverse split on the way into shared memory (K4F); K2/K4 at 4 CTAs per SM with the smaller twiddle working set (R4)
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
    print('$name'.ljust(20), 'value', round(d['value']), 'ms/block', round(d['ms_per_step']/k,4), 'stepfrac', round(d['step_roofline']['frac'],3), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), {a: round(b,4) for a,b in d['step_roofline']['kernels_ms'].items()})
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
AW_LIBRARY=$L/libairwave_K4F.so timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 2>&1 | tail -1
AW_LIBRARY=$L/libairwave_K4F.so AW_FUSED_TILE=0 timeout 600 python -m pytest tests/test_gpu_convolution.py tests/test_gpu_eq.py -m gpu -q --timeout 300 2>&1 | tail -1
for v in BASE K4F R4 BASE K4F R4; do b C5-4096_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload C5-4096; done
for v in BASE K4F; do b C2s_$v AW_LIBRARY=$L/libairwave_$v.so AW_FUSED_TILE=0 -- --workload C2; done
