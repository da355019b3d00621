#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "2gpu rc=$?"; tail -1 gpurun_out/bench_2gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','gpu_launches','clocks','host_affinity')}, d['e2e']['value'], d['roofline']['frac'])"
nvidia-smi topo -m | head -12
