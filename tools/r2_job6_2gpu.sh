#!/bin/bash
# 2-GPU box: sharding proof, concurrent PCIe floor at N=1,2, bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 500 2>&1 | tail -5
for N in 1 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N tools/pcie_scaling.py 2>&1 | grep "^N=" | tee -a gpurun_out/pcie_scaling_2gpu.txt
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 --e2e-steps 40 2>gpurun_out/b_2gpu.err | tail -1 > gpurun_out/b_2gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_2gpu.json'))
print('N=2 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'sharding', d.get('sharding'), 'single', (d.get('single_block_calls') or {}).get('value'))
PY
timeout 300 python bench.py --gpus 1 --steps 200 --warmup 20 --e2e-steps 40 --no-cpu 2>gpurun_out/b_1gpu.err | tail -1 > gpurun_out/b_1gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_1gpu.json'))
print('N=1 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'single', (d.get('single_block_calls') or {}).get('value'), d['latency_ms'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload C5-offline --offline-seconds 10 --offline-blocks 64,256,1024 2>gpurun_out/b_off2.err | tail -1 > gpurun_out/b_off2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_off2.json'))
for e in d['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items() if k in ('block','e2e_value','device_value_per_gpu','roofline_frac','d2h_gbs_per_gpu')})
PY
