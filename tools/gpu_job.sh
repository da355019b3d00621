#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dbg_twins2.py 2>&1 | grep -v "^   stream" | tail -10
PYTEST_TIMEOUT=1800 PYTEST_ARGS="--timeout 900" bash tools/gpu_check.sh
bash tools/bench_all.sh C2 C3 C5-512
