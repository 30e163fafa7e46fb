#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/step_n.py 64 3 > gpurun_out/plain_ncu.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 23 -c 1 -o gpurun_out/prof_conv8 -f python scripts/step_n.py 64 3 > gpurun_out/ncu_conv8.log 2>&1; echo "ncu8 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 1 -o gpurun_out/prof_conv3 -f python scripts/step_n.py 64 3 > gpurun_out/ncu_conv3.log 2>&1; echo "ncu3 exit $?"
ls -la gpurun_out/*.ncu-rep
