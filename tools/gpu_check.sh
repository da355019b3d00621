#!/bin/bash
# GPU-side helper (run under gpurun): smoke + the GPU test-suite, each under its own timeout so a hung kernel cannot hold the box.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q --timeout 240 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
