#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|^ok|Invalid|Error" gpurun_out/sanitize_memcheck.log | tail -8
timeout 1500 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize.py > gpurun_out/sanitize_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|Barrier error|Error" gpurun_out/sanitize_synccheck.log | head -5
timeout 1500 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize.py > gpurun_out/sanitize_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|Uninitialized|Error" gpurun_out/sanitize_initcheck.log | head -10
