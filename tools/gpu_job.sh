#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 30 python tools/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "^ok|RACECHECK SUMMARY|ERROR SUMMARY|Race reported|hazard" gpurun_out/sanitize_racecheck.log | head -40
