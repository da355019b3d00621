#!/bin/bash
# radix-16 twiddles as six table reads + nine products (TP) against fifteen reads (N1)
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
AW_LIBRARY=$L/libairwave_TP.so timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 2>&1 | tail -1
for w in C5-4096 C5-2048 C5-1024 C2 C5-4096 C5-2048 C5-1024 C2; do for v in N1 TP; do b ${w}_${v} AW_LIBRARY=$L/libairwave_$v.so -- --workload $w; done; done
