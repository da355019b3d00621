#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_convolution.py -m gpu -x -q -k "many_profile or per_range or independent" --timeout 600 2>&1 | tail -3
bash tools/bench_all.sh F3 C2
