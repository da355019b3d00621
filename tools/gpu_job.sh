#!/bin/bash
PYTEST_TIMEOUT=600 bash tools/gpu_check.sh
timeout 300 bash tools/bench_all.sh F3 C4
