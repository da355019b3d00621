#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 40 > gpurun_out/bench_2gpu.log 2>&1; echo "2gpu rc=$?"; tail -1 gpurun_out/bench_2gpu.log | cut -c1-1800
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_ref_2gpu.log 2>&1; echo "ref 2gpu rc=$?"; tail -1 gpurun_out/bench_ref_2gpu.log | cut -c1-900
timeout 600 python bench.py --steps 200 --warmup 40 > gpurun_out/bench_1gpu.log 2>&1; echo "1gpu rc=$?"; tail -1 gpurun_out/bench_1gpu.log | cut -c1-2500
