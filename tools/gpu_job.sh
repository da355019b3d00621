#!/bin/bash
mkdir -p gpurun_out
BENCH_ENV="AW_PERSISTENT_DEBUG=3" bash tools/bench_all.sh C5-1024 C5-2048 C2 C5-64
BENCH_ENV="AW_PERSISTENT_DEBUG=1" bash tools/bench_all.sh C5-1024 C5-2048
BENCH_ENV="AW_PERSISTENT_DEBUG=2" bash tools/bench_all.sh C5-1024 C5-2048
