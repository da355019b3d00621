#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/tmabw tools/tmabw.cu && timeout 300 /tmp/tmabw > gpurun_out/tmabw.log 2>&1; echo rc=$?
cat gpurun_out/tmabw.log
