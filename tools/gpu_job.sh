#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "paths_agree or small_and_ragged" --timeout 300 > gpurun_out/pytest_paths.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_paths.log
