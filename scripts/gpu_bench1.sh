#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r1.json; tail -5 gpurun_out/bench_r1.err
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 300 -c 3 -o gpurun_out/prof_conv_r1 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | head -30
