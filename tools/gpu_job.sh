#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "paths_agree or c2_full" --timeout 600 > gpurun_out/pytest_paths.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_paths.log
bash tools/bench_all.sh C2 C3 C4 C5-512
