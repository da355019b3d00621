#!/bin/bash
mkdir -p gpurun_out
BENCH_ENV="AW_PERSISTENT_DEBUG=16" bash tools/bench_all.sh C5-512 C2
BENCH_ENV="AW_PERSISTENT_DEBUG=32" bash tools/bench_all.sh C5-512 C2
