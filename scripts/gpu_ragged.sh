#!/bin/bash
# Ragged (non-power-of-two) maps on the tensor-core path: kernel cases, then the C1 model, each under its own timeout.
mkdir -p gpurun_out
{
  timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 --timeout-method=thread -k "conv_paths and (28 or 14-14 or 7-7 or 12-20)" 2>&1 | tail -40
  echo "--- rejects + model"
  timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --timeout 120 --timeout-method=thread -k "rejects or eps_bf16 or untileable" -s 2>&1 | tail -40
} > gpurun_out/ragged.log 2>&1
tail -60 gpurun_out/ragged.log
