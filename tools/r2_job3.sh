#!/bin/bash
mkdir -p gpurun_out
b() { # name, dir, env..., -- args
  name=$1; dir=$2; shift; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (cd $dir && env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-300} --warmup 40 --no-cpu --e2e-steps 3 "$@" 2>$OLDPWD/gpurun_out/b_$name.err | tail -1 > $OLDPWD/gpurun_out/b_$name.json)
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    mb=d.get('multiblock') or {}
    print('$name', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['step_roofline']['frac'],3), 'p99', round(d['latency_ms']['p99'],4), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'mb', round(mb.get('ms_per_block',0),4), round(mb.get('frac_of_peak_by_per_block_algorithmic_bytes',0),3))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
VB=$PWD/airwave_b200/lib/libairwave_varB.so
for w in C2 C5-512 C3; do
  b ${w}_old _r1_baseline AW_X=0 -- --workload $w
  b ${w}_new . AW_X=0 -- --workload $w
  b ${w}_new_ring0 . AW_KP_RING_EXTRA=0 -- --workload $w
  b ${w}_varB . AW_LIBRARY=$VB -- --workload $w
  b ${w}_varB_ring0 . AW_LIBRARY=$VB AW_KP_RING_EXTRA=0 -- --workload $w
  b ${w}_old2 _r1_baseline AW_X=0 -- --workload $w
done
timeout 300 python bench.py --workload C1 --steps 200 --warmup 20 --no-cpu --e2e-steps 3 2>gpurun_out/b_c1.err | tail -1 > gpurun_out/b_c1.json; python -c "
import json; d=json.load(open('gpurun_out/b_c1.json')); print(json.dumps(d.get('sync_latency'), indent=1)[:3000])" || tail -5 gpurun_out/b_c1.err
timeout 300 python bench.py --workload C5-offline --offline-seconds 5 --offline-blocks 64,256,1024,4096 2>gpurun_out/b_off.err | tail -1 > gpurun_out/b_off.json; python -c "
import json; d=json.load(open('gpurun_out/b_off.json'))
for e in d['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items()})" || tail -5 gpurun_out/b_off.err
