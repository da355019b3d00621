#!/bin/bash
mkdir -p gpurun_out
b() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --steps ${STEPS:-200} --warmup 20 --no-cpu --e2e-steps 3 --multiblock 0 "$@" 2>gpurun_out/b_$name.err | tail -1 > gpurun_out/b_$name.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b_$name.json'))
    k=d['config']['blocks_per_step']
    print('$name'.ljust(28), 'ms/block', round(d['ms_per_step']/k,4), 'frac', round(d['step_roofline']['frac'],3), d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'))
except Exception as e:
    print('$name', 'FAILED', e, open('gpurun_out/b_$name.err').read()[-400:])
PY
}
for keep in 0 25 50 75 100; do b c2_k4_keep$keep AW_KP_KEEP=$keep -- --workload C2 --blocks-per-call 4; done
b c2_k4_bm AW_KP_ORDER=0 -- --workload C2 --blocks-per-call 4
for keep in 50 100; do b c2_k4_t2_keep$keep AW_KP_KEEP=$keep AW_PERSISTENT_TILE=2 -- --workload C2 --blocks-per-call 4; done
for k in 2 8 16; do b c2_k${k} AW_X=0 -- --workload C2 --blocks-per-call $k --e2e-frames 4096; done
for keep in 0 50 100; do b c564_k16_keep$keep AW_KP_KEEP=$keep -- --workload C5-64 --blocks-per-call 16 --e2e-frames 4096; done
b c564_k64 AW_X=0 -- --workload C5-64 --blocks-per-call 64 --e2e-frames 4096
b c564_k64_t2 AW_PERSISTENT_TILE=2 AW_KP_KEEP=100 -- --workload C5-64 --blocks-per-call 64 --e2e-frames 4096
b c5128_k32 AW_X=0 -- --workload C5-128 --blocks-per-call 32 --e2e-frames 4096
b c5512_k8 AW_X=0 -- --workload C5-512 --blocks-per-call 8 --e2e-frames 4096
b c5512_k8_keep100 AW_KP_KEEP=100 -- --workload C5-512 --blocks-per-call 8 --e2e-frames 4096
b c51024_k4 AW_X=0 -- --workload C5-1024 --blocks-per-call 4 --e2e-frames 4096
b c52048_k2 AW_X=0 -- --workload C5-2048 --blocks-per-call 2 --e2e-frames 4096
b c3_k4 AW_X=0 -- --workload C3 --blocks-per-call 4 --e2e-frames 2048
b c3_k4_keep100 AW_KP_KEEP=100 -- --workload C3 --blocks-per-call 4 --e2e-frames 2048
b c4_k4 AW_X=0 -- --workload C4 --blocks-per-call 4
