#!/bin/bash
# GPU-side helper (run under gpurun): sweep the execution plan of the block kernels on the bench workload.
#   AW_FUSED_TILE = 4/2/1 forces the fused K2+K3+K4 kernel with that many streams per CTA, 0 forces the split kernels
#   AW_MAC_TILE   = streams per thread of the stand-alone K3
WORKLOAD=${1:-C2}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" python bench.py --workload $WORKLOAD --steps 100 --warmup 20 --no-cpu --e2e-steps 3 2>&1 | tail -1 > gpurun_out/sweep_${WORKLOAD}_$tag.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sweep_${WORKLOAD}_$tag.json'))
    print('$tag', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'step_frac', round(d['step_roofline']['frac'],3), d['config']['plan'], {k: round(v,4) for k,v in d['step_roofline']['kernels_ms'].items()})
except Exception as e:
    print('$tag', 'FAILED', e, open('gpurun_out/sweep_${WORKLOAD}_$tag.json').read()[-400:])
PY
}
run fused4 AW_FUSED_TILE=4
run fused2 AW_FUSED_TILE=2
run fused1 AW_FUSED_TILE=1
run split_mac2 AW_FUSED_TILE=0 AW_MAC_TILE=2
run split_mac4 AW_FUSED_TILE=0 AW_MAC_TILE=4
run auto AW_DUMMY=1
run fused4_nostagger AW_FUSED_TILE=4 AW_FUSED_STAGGER=0
run fused2_nostagger AW_FUSED_TILE=2 AW_FUSED_STAGGER=0
