#!/bin/bash
PYTEST_TIMEOUT=1800 PYTEST_ARGS="--timeout 900" bash tools/gpu_check.sh
