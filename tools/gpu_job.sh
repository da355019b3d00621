#!/bin/bash
PYTEST_TIMEOUT=1800 PYTEST_ARGS="--timeout 900" bash tools/gpu_check.sh
bash tools/bench_all.sh C5-4096
BENCH_ENV="AW_FUSED_TILE=0" bash tools/bench_all.sh C5-2048 C5-1024 C2
