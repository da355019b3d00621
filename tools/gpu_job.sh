#!/bin/bash
mkdir -p gpurun_out
PYTEST_TIMEOUT=1800 PYTEST_ARGS="--timeout 900" bash tools/gpu_check.sh
bash tools/bench_all.sh C4 C2 C5-512
BENCH_ENV="AW_EQ_FUSION=0" bash tools/bench_all.sh C4
