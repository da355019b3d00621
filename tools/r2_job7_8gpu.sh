#!/bin/bash
# 8-GPU box: PCIe floor at N=1,2,4,8; sharding proof at 4 and 8; C2 bench at N=8; C5-offline (60 s x 2048 streams/GPU, block sweep) at N=8
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
lscpu | egrep "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/host_8gpu.txt 2>&1; free -g >> gpurun_out/host_8gpu.txt
rm -f gpurun_out/pcie_scaling_8gpu.txt
for N in 1 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$N tools/pcie_scaling.py 2>&1 | grep "^N=" | tee -a gpurun_out/pcie_scaling_8gpu.txt
done
timeout 400 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 350 -k "4 or 8" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 500 --warmup 20 --e2e-steps 40 2>gpurun_out/b_8gpu.err | tail -1 > gpurun_out/b_8gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_8gpu.json'))
print('N=8 value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'sharding', d.get('sharding'), 'single', (d.get('single_block_calls') or {}).get('value'), d['clocks'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 4 --steps 200 --warmup 20 --e2e-steps 40 --no-single-block 2>gpurun_out/b_4gpu.err | tail -1 > gpurun_out/b_4gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_4gpu.json'))
print('N=4 value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --workload C5-offline --offline-seconds 60 2>gpurun_out/b_off8.err | tail -1 > gpurun_out/b_off8.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_off8.json'))
for e in d['sweep']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items() if k in ('block','e2e_value','elapsed_s','device_value_per_gpu','roofline_frac','d2h_gbs_per_gpu','device_ms_per_block')})
PY
