#!/bin/bash
# GPU-side (under gpurun): first round-2 check of KP v2 (multi-block launches, spare ring slot, tail tiles)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
b() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps 400 --warmup 40 --no-cpu --e2e-steps 3 "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    print('$name', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['step_roofline']['frac'],3), 'p99', round(d['latency_ms']['p99'],4), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
b c2_k1 AW_X=0 -- --workload C2
b c2_k4 AW_X=0 -- --workload C2 --blocks-per-call 4
b c2_k4_bm AW_KP_ORDER=0 -- --workload C2 --blocks-per-call 4
b c2_k4_keep0 AW_KP_KEEP=0 -- --workload C2 --blocks-per-call 4
b c2_k4_keep100 AW_KP_KEEP=100 -- --workload C2 --blocks-per-call 4
b c2_k4_t2_keep100 AW_KP_KEEP=100 AW_PERSISTENT_TILE=2 -- --workload C2 --blocks-per-call 4
b c2_k4_t2_keep50 AW_KP_KEEP=50 AW_PERSISTENT_TILE=2 -- --workload C2 --blocks-per-call 4
b c2_k4_t2_bm AW_KP_ORDER=0 AW_PERSISTENT_TILE=2 -- --workload C2 --blocks-per-call 4
b c564_k1 AW_X=0 -- --workload C5-64
b c564_k16 AW_X=0 -- --workload C5-64 --blocks-per-call 16
b c564_k16_keep100 AW_KP_KEEP=100 -- --workload C5-64 --blocks-per-call 16
b c5128_k1 AW_X=0 -- --workload C5-128
b c5512_k1 AW_X=0 -- --workload C5-512
b c5512_k2 AW_X=0 -- --workload C5-512 --blocks-per-call 2
b c51024_k1 AW_X=0 -- --workload C5-1024
b c52048_k1 AW_X=0 -- --workload C5-2048
b c3_k1 AW_X=0 -- --workload C3
b c4_k1 AW_X=0 -- --workload C4
b c4_k4 AW_X=0 -- --workload C4 --blocks-per-call 4
