#!/bin/bash
# final validation of the round: smoke + GPU suite, the driver's bench command for both arms, the offline sweep
bash tools/gpu_check.sh
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02_bench_C2_driver_cmd.err | tail -1 > gpurun_out/r02_bench_C2_driver_cmd.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_C2_driver_cmd.json'));print('ours', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['clocks'], 'cpu', d['cpu_baseline']['value'])"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>gpurun_out/r02_bench_reference.err | tail -1 > gpurun_out/r02_bench_reference.json
python -c "import json;d=json.load(open('gpurun_out/r02_bench_reference.json'));print('reference', round(d['value']), d['cpu_baseline']['cores'])"
timeout 600 python bench.py --workload C5-offline --steps 40 --warmup 3 2>gpurun_out/r02_offline_1gpu.err | tail -1 > gpurun_out/r02_offline_1gpu.json
python -c "
import json;d=json.load(open('gpurun_out/r02_offline_1gpu.json'))
for r in d['sweep']: print(r['block'], 'e2e', round(r['e2e_value']), 'device', round(r['device_value_per_gpu']), 'frac', round(r['roofline_frac'],3), 'd2h GB/s', round(r['d2h_gbs_per_gpu'],1))
print(d['clocks'])"
