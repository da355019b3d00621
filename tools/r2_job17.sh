#!/bin/bash
mkdir -p gpurun_out
b() { name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-300} --warmup 20 --no-cpu --e2e-steps 3 --no-single-block "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config'].get('blocks_per_step',1)
    print('$name'.ljust(20), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), {a: round(b,4) for a,b in d['step_roofline']['kernels_ms'].items()})
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
L=$PWD/airwave_b200/lib
b C5-4096_base AW_X=0 -- --workload C5-4096
for v in SA4 SA5 SA6; do b C5-4096_$v AW_LIBRARY=$L/libairwave_$v.so -- --workload C5-4096; done
b C5-4096_base2 AW_X=0 -- --workload C5-4096
for v in SA4 SA5; do AW_LIBRARY=$L/libairwave_$v.so timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 -k "4096" 2>&1 | tail -1; done
