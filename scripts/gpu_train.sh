#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/pytest_train.log 2>&1; echo "exit $?"; grep -E "passed|failed|Error|error|assert|worst" gpurun_out/pytest_train.log | head -40
