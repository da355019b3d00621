#!/bin/bash
# GPU-side: the round's final single-GPU evidence — full test-suite (with the C4 chain error record), sanitizer sweep, the default
# bench line as the driver runs it (20 steps) and sustained (2000 steps), C1 with the synchronous-latency figures, the reference
# arm, the 60 s offline sweep on one GPU, and the e2e path with 4096-frame submits.
mkdir -p gpurun_out
AW_EVIDENCE_DIR=gpurun_out timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 2000 --warmup 40 2>gpurun_out/r02_bench_C2_sustained.err | tail -1 > gpurun_out/r02_bench_C2_sustained.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02_bench_C2_driver_cmd.err | tail -1 > gpurun_out/r02_bench_C2_driver_cmd.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_reference.json
timeout 600 python bench.py --workload C1 --steps 400 --warmup 40 --no-cpu 2>gpurun_out/r02_bench_C1.err | tail -1 > gpurun_out/r02_bench_C1.json
timeout 600 python bench.py --workload C5-offline --offline-seconds 60 2>gpurun_out/r02_offline_1gpu.err | tail -1 > gpurun_out/r02_offline_1gpu.json
timeout 600 python bench.py --steps 400 --warmup 40 --no-cpu --e2e-frames 4096 --e2e-steps 20 --no-single-block 2>/dev/null | tail -1 > gpurun_out/r02_bench_C2_e2e4096.json
python - <<'PY'
import json
for n in ['r02_bench_C2_sustained','r02_bench_C2_driver_cmd','r02_bench_reference','r02_bench_C1','r02_bench_C2_e2e4096']:
    try:
        d=json.load(open(f'gpurun_out/{n}.json'))
        print(n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round((d.get('roofline') or {}).get('frac',0),3), (d.get('single_block_calls') or {}).get('value'), d.get('clocks'))
    except Exception as e:
        print(n, 'FAILED', e)
d=json.load(open('gpurun_out/r02_offline_1gpu.json'))
for e in d['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items() if k in ('block','e2e_value','device_value_per_gpu','roofline_frac','d2h_gbs_per_gpu')})
PY
