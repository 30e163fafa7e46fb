#!/bin/bash
mkdir -p gpurun_out
{
  timeout 400 python -m pytest tests/test_gpu_chain_full.py -m gpu -q --timeout 200 --timeout-method=thread -k "c1_full" -s 2>&1 | tail -12
  timeout 200 python scripts/c1_time.py 2>&1 | tail -4
} > gpurun_out/ragged2.log 2>&1
tail -30 gpurun_out/ragged2.log
