#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_kp2.py 2>&1 | grep -B1 "bad blocks \[[0-9]" | tail -20; echo DBG-END
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
b() { # name, dir, env..., -- args
  name=$1; dir=$2; shift; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (cd $dir && env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-300} --warmup 40 --no-cpu --e2e-steps 3 "$@" 2>$OLDPWD/gpurun_out/b_$name.err | tail -1 > $OLDPWD/gpurun_out/b_$name.json)
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    mb=d.get('multiblock') or {}
    print('$name'.ljust(22), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'mb', round(mb.get('ms_per_block',0),4), round(mb.get('frac_of_peak_by_per_block_algorithmic_bytes',0),3), d['config']['plan'].get('tensor_map_tma'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
for w in C2 C3 C5-64 C5-128 C5-512 C5-1024 C5-2048 C4; do
  b ${w}_old _r1_baseline AW_X=0 -- --workload $w
  b ${w}_tma . AW_X=0 -- --workload $w
  b ${w}_bulk . AW_KP_TENSOR_TMA=0 -- --workload $w
done
