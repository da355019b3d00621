#!/bin/bash
PYTEST_TIMEOUT=1800 PYTEST_ARGS="--timeout 900" bash tools/gpu_check.sh
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py 2>&1 | grep -E "ERROR SUMMARY|1024|2048"
