#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
b() { name=$1; dir=$2; shift; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (cd $dir && env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-300} --warmup 20 --no-cpu --e2e-steps 3 "$@" 2>$OLDPWD/gpurun_out/b_$name.err | tail -1 > $OLDPWD/gpurun_out/b_$name.json)
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config'].get('blocks_per_step',1); sb=d.get('single_block_calls') or {}
    print('$name'.ljust(20), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), 'single', round(sb.get('ms_per_block',0),4), round(sb.get('frac',0),3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
for w in C5-1024 C5-2048; do
  b ${w}_prev _prev_build AW_X=0 -- --workload $w
  b ${w}_pooled . AW_X=0 -- --workload $w
  b ${w}_prev_k4 _prev_build AW_X=0 -- --workload $w --e2e-frames 4096
  b ${w}_pooled_k4 . AW_X=0 -- --workload $w --e2e-frames 4096
done
python tools/stress_round2.py > gpurun_out/r02_stress.txt 2>&1; tail -2 gpurun_out/r02_stress.txt
