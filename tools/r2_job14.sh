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
    print('$name'.ljust(20), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), d['step_roofline']['kernels_ms'])
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
L=$PWD/airwave_b200/lib
for v in E32a E32b; do
  AW_LIBRARY=$L/libairwave_$v.so timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -q --timeout 300 -k "1024 or 2048 or 4096" 2>&1 | tail -2
done
for w in C5-1024 C5-2048 C5-4096; do
  b ${w}_base AW_X=0 -- --workload $w
  b ${w}_E32a AW_LIBRARY=$L/libairwave_E32a.so -- --workload $w
  b ${w}_E32b AW_LIBRARY=$L/libairwave_E32b.so -- --workload $w
done
