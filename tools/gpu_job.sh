#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_eq.py -m gpu -x -q --timeout 60 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "c4 or fused_eq" --timeout 120 2>&1 | tail -2
timeout 300 bash tools/bench_all.sh C4 F3
