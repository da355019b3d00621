#!/bin/bash
bash tools/bench_all.sh F3
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "many_profile or c2_full" --timeout 600 2>&1 | tail -2
